set -x
cd $GRAFT_REPO_ROOT
for v in "" _g2 _g3; do
  export CINTB200_LIB=$GRAFT_REPO_ROOT/libcint_b200/libcint_b200$v.so
  echo "== variant '$v'"
  for args in "1 1 0 0 2" "2 2 2 2 2 640" "3 3 3 3 2 100" "2 1 1 0 2" "5 5 5 5 1 8"; do timeout 120 python tools/quick_sweep1.py $args 2>/dev/null | tail -1 | cut -c1-40,100-; done
  timeout 300 python tools/quick_ip1.py 2>/dev/null | tail -1
  timeout 300 python tools/time_variant.py 8 c2h6_ccpvqz | cut -c1-200
done
