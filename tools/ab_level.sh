# Development tool: one pyramid level (C, H, W; 40 frames, channels-last) forward + backward for the default library and
# every variant under c2m_b200/variants.  Output: gpurun_out/${TAG}_level.txt
T=${TAG:-ab}; C=${1:-32}; H=${2:-256}; W=${3:-512}
OUT=gpurun_out/${T}_level.txt
: > $OUT
for rep in 1 2; do
for lib in "" $(ls c2m_b200/variants/*.so 2>/dev/null); do
  C2M_WARP_LIB=${lib:+$PWD/$lib} python - $C $H $W "${lib:-default}" >> $OUT <<'PY'
import sys, torch, c2m_b200
C, H, W = map(int, sys.argv[1:4]); N = 40; dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(N, C, H, W, device=dev, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
ii = torch.arange(H, device=dev).view(1, H, 1).float(); jj = torch.arange(W, device=dev).view(1, 1, W).float()
fx = 8 * torch.sin(6.2832 * ii / (H / 2)) * torch.cos(6.2832 * jj / (W / 2)); fy = 8 * torch.cos(6.2832 * ii / (H / 2)) * torch.sin(6.2832 * jj / (W / 2))
flow = (torch.stack([fx.expand(N, H, W), fy.expand(N, H, W)], 1) + torch.randn(N, 2, H, W, device=dev, generator=g)).requires_grad_(True)
mask = torch.sigmoid(torch.randn(N, 1, H, W, device=dev, generator=g)).requires_grad_(True)
gout = torch.randn(N, C, H, W, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
evs = []
def step(rec=False):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); o = c2m_b200.warp_blend(x, flow, mask); e[1].record()
    torch.autograd.grad(o, [x, flow, mask], gout); e[2].record()
    if rec: evs.append(e)
for _ in range(5): step()
for _ in range(20): step(True)
torch.cuda.synchronize()
f = sum(e[0].elapsed_time(e[1]) for e in evs) / len(evs); b = sum(e[1].elapsed_time(e[2]) for e in evs) / len(evs)
print("%-34s C=%d %dx%d  fwd %.4f ms  bwd %.4f ms  step %.4f ms" % (sys.argv[4], C, H, W, f, b, f + b))
PY
done
done
cat $OUT
