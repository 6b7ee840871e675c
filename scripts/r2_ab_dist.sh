# A/B of library variants: the C3 step (scripts/r2_ab.sh) and the distance query over its pairs
bash scripts/r2_ab.sh
for lib in physkit_b200/libpk_collide.so physkit_b200/libpk_v_*.so; do
  echo "distance $(basename $lib .so): $(PK_COLLIDE_LIB=$PWD/$lib PK_C4_PAIRS=0 timeout 300 python scripts/r2_distance.py 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms"],3), "ms", d["separated"])')"
done
