#!/bin/bash
# one-knob variant builds for the per-class tuning (tools/tune_classes.py): only the kernel instantiation objects are rebuilt.
# Round-2 runs: va = reg minb 3 / coop (3,2,2); vb = reg minb 4 / coop (5,4,3); vc = reg unroll 2 / coop (6,5,4); second pass
# (register kernels only, against the tuned defaults): vd = minb 3 + unroll 2, ve = minb 4 + unroll 2.
cd /root/repo/libcint_b200/csrc
build() {  # name, flags, objects to rebuild
  rm -rf build_$1; mkdir -p build_$1; cp -p build/*.o build_$1/; rm -f build_$1/kern_reg_inst_*.o; [ "$3" = reg ] || rm -f build_$1/kern_coop_inst_*.o
  make -j8 VARIANT=_$1 EXTRA="-DTUNE_OFF $2" > build_$1/make.log 2>&1 || { echo "variant $1 FAILED"; tail -5 build_$1/make.log; }
}
for v in "$@"; do
  case $v in
    va) build va "-DREG_MIN_BLOCKS=3 -DCOOP_MB16=3 -DCOOP_MB32=2" ;;
    vb) build vb "-DREG_MIN_BLOCKS=4 -DCOOP_MB16=5 -DCOOP_MB32=4 -DCOOP_MBX=3" ;;
    vc) build vc "-DREG_UQ_UNROLL=2 -DCOOP_MB16=6 -DCOOP_MB32=5 -DCOOP_MBX=4" ;;
    vd) build vd "-DREG_MIN_BLOCKS=3 -DREG_UQ_UNROLL=2" reg ;;
    ve) build ve "-DREG_MIN_BLOCKS=4 -DREG_UQ_UNROLL=2" reg ;;
  esac
done
ls -la ../*.so
