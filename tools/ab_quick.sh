# Development tool: step / forward / backward / gather times of the default library and of every variant under
# c2m_b200/variants (tools/build_variants.py), two runs each.  Output: gpurun_out/${TAG}_ab.txt
T=${TAG:-ab}
OUT=gpurun_out/${T}_ab.txt
: > $OUT
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --e2e-steps 0 --torch-cuda-steps 0 --no-pyramids --no-configs --no-other-layout"
for rep in 1 2; do
for lib in "" $(ls c2m_b200/variants/*.so 2>/dev/null); do
    name=${lib:-default}
    C2M_WARP_LIB=${lib:+$PWD/$lib} $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-36s step %.4f ms  fwd %.4f  bwd %.4f  gather %.4f  (frac %.3f, step frac %.3f)' % ('$name', d['ms_per_step'], r['fwd']['ms'], r['bwd']['ms'], r['kernels_alone']['bwd_gather_ms'], r['frac'], r['fwd_bwd']['frac']))" >> $OUT
done
done
cat $OUT
