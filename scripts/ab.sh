# A/B aid: time the chain launch with each dppo_b200/lib/variant_<X>.so in turn (VARIANTS="A B"), same box, same process setup
cd dppo_b200/lib
cp libdppo_b200.so keep.so
for v in $VARIANTS; do cp variant_$v.so libdppo_b200.so; echo "variant $v"; (cd ../..; timeout 100 python scripts/${AB_SCRIPT:-ab_time.py} ${AB_ARGS:-walker2d 200} 2>&1 | tail -${AB_TAIL:-1}); done
cp keep.so libdppo_b200.so; rm keep.so
