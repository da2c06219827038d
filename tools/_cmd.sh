timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2p_pytest.log
python - > gpurun_out/r2p_occ.log 2>&1 <<'PY'
import torch, c2m_b200, sys
sys.path.insert(0,'tests')
from test_occlusion_map import _torch_reference
dev=torch.device('cuda',0)
flow=torch.randn(40,2,256,512,device=dev)*3
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print('ours   get_occlusion_map 40x256x512: %.3f ms'%t(lambda: c2m_b200.get_occlusion_map(flow)))
print('torch  composition             : %.3f ms'%t(lambda: _torch_reference(flow).clamp(0,1)))
PY
