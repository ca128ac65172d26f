cd dppo_b200/lib
cp libdppo_b200.so keep.so
for v in $VARIANTS; do cp variant_$v.so libdppo_b200.so; echo "variant $v"; (cd ../..; timeout 100 python scripts/ab_time.py walker2d 200 2>&1 | tail -1); done
cp keep.so libdppo_b200.so
