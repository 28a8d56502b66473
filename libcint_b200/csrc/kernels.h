// Launch-side interface between the host engine (engine.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include "types.h"

struct GenericClass {           // uniform per launch of eri_generic_kernel
    int la, lb, lc, ld;
    int nroots;
    int nreff;                  // quadrature points actually summed (2*nroots for short-range Coulomb)
    int ncab, nccd;             // contraction combinations of the bra / ket pair
    int nE, nF;                 // Cartesian components of [e0| (e = la..la+lb) and |f0]
    int acc_in_smem, work_in_smem;
    int work_size;              // doubles per ping-pong buffer of the epilogue
    int c2s_off[4];             // offsets of the four c2s matrices inside EngineParams::c2s
    double *scratch;            // global scratch, scratch_per_block doubles per block
    size_t scratch_per_block;
    int map_ints;               // ints of shared memory for the per-level HRR map (largest level of the class)
    int csplit;                 // threads sharing the contraction-combination updates of one [e0|f0] component (power of two)
    int pbatch;                 // primitive quartets processed per batch (root / recurrence phases run for all of them at once)
    int wide;                   // the quadrature runs on the wide kernel (kern_wide.cu), this kernel does the epilogue only
    int epilogue_only;          // set per launch: accumulators of task task_base + blockIdx.x are already in the block's scratch
    long long task_base;
};

struct GenericLaunch { int grid, threads; size_t smem; };

int generic_setup_constants();
int generic_plan(GenericClass *C, GenericLaunch *L, int la, int lb, int lc, int ld, int ncab, int nccd,
                 int cart, long long ntasks, const int *c2s_off_table, int short_range = 0);
int generic_launch(const EngineParams &P, const GenericClass &C, const GenericLaunch &L, const Task *tasks,
                   long long ntasks, double *out, int *nonzero, unsigned long long *counters, cudaStream_t stream,
                   const TileParams *tile = nullptr, const long long *uprefix = nullptr);

// Grid of a persistent tile-kernel launch (engine.cu): raises the function's dynamic shared-memory limit when needed (fails with
// -1 if `smem` exceeds what the device offers -- callers route such classes to the generic kernel) and returns
// min(total work items, blocks per SM x number of SMs).  Blocks per SM: exactly the resident capacity of the function (every
// block stages its Rys table once and then pulls work items from the launch's counter); CINTB200_PBLOCKS=n forces n per SM.
int tile_grid_blocks(const void *fn, int threads, size_t smem, long long total);
int tile_smem_limit();                        // largest dynamic shared memory per block the device allows (opt-in), bytes

// wide kernel (kern_wide.cu): high-l classes without a register / cooperative instantiation
int wide_eligible(int la, int lb, int lc, int ld, int ncab, int nccd, int short_range);
int wide_launch(const EngineParams &P, const GenericClass &C, const Task *tasks, long long task_base, long long ntasks, int ntask_here,
                int *nonzero, cudaStream_t stream, const TileParams *tile);

// register kernels (kern_reg_inst*.cu): thread per quartet, compile-time class
typedef void (*RegKernelFn)(const TileParams);
RegKernelFn reg_kernel_lookup(int la, int lb, int lc, int ld, int nct, int ncu, int rs = 0, int cart = 0);     // rs: range-separated variant, cart: Cartesian output
int rys_tab_nint(int nroots);
int rys_fast_nint(int nroots);
int rys_fast_off(int nroots);
int reg_kernel_launch(RegKernelFn fn, int nroots, int ncu, const TileParams &P, int grid_x, int grid_y, cudaStream_t stream);
size_t reg_kernel_smem(RegKernelFn fn, int nroots, int ncu, int umax);       // dynamic shared memory such a launch needs

// cooperative kernels (kern_coop_inst*.cu): FS lanes per quartet
struct CoopInfo { int fs, xsz, nroots; };
RegKernelFn coop_kernel_lookup(int tla, int tlb, int ula, int ulb, int nct, int ncu, CoopInfo *info, int rs = 0, int cart = 0);
int coop_kernel_launch(RegKernelFn fn, const CoopInfo &info, int ncu, const TileParams &P, int grid_x, int grid_y, cudaStream_t stream);
size_t coop_kernel_smem(const CoopInfo &info, int ncu, int umax);
