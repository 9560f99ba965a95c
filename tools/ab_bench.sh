L=ml_conformer_generator_b200/libmlcg_b200.so
cp $L /tmp/orig.so
for i in 1 2; do for v in A B; do cp build_ab/$v.so $L; timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v', round(d['value'], 1), 'edge_ms', round(d['roofline']['launch_ms'], 3), d['clocks']['sm_mhz'])"; done; done
cp /tmp/orig.so $L
