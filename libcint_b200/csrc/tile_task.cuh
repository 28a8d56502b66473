// Tile mode of the block-per-quartet kernels (kern_generic.cu, kern_wide.cu): work item -> shell quartet of a whole-job tile.
#pragma once
#include "types.h"

// tile mode: work item w -> (u, t) of the rectangle (this rank's kets) x (T pairs of the chunk);
// returns bra = -1 for quartets outside the reference loop bound k <= i
__device__ __forceinline__ Task tile_task(const TileParams &T, long long w)
{
    const int nT = T.t_end - T.t_begin;
    const int j = (int)(w / nT);
    const int t = T.t_begin + (int)(w - (long long)j * nT);
    const int u = T.u_first + T.u_step * j;
    Task k;
    k.bra = (T.tri && T.tI[t] < T.uK[u]) ? -1 : T.tpair[t];      // tri = 1 lists are shell-sorted; a predicate is enough here
    k.ket = T.upair[u];
    k.sa = T.tstride[t];
    k.sb = T.tstride[T.NT + t];
    k.sc = (long long)T.ustride[u] * T.ld;
    k.sd = (long long)T.ustride[T.NU_all + u] * T.ld;
    k.off = (T.trow[t] - T.row0) + T.ucol[u] * T.ld;
    k.flags = 0; k.pad = 0;
    return k;
}

