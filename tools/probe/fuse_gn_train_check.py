"""One training step with the fused conv+GroupNorm kernels in the keep-raw forward (engine.fuse_gn_train) vs the default
conv + GroupNorm launches: per-tensor gradient differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden"))
import torch
from weights import make_state_dict, synthetic_chirps, gaussian
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import CustomDiffusion, UNet1D
from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import FusedTrainStep

def run(fuse, B, L, in_ch, cc, steps=1, graph=False):
    sd = make_state_dict(in_ch=in_ch, cond_in_ch=cc, seed=2)
    m = UNet1D(in_ch=in_ch, cond_in_ch=cc, use_selfcond=True, compute_dtype="bf16")
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    st = FusedTrainStep(m, CustomDiffusion(T=1000, device="cuda"), B, L, compute_dtype="bf16", p_uncond=0.2, warmup_steps=10, total_steps=100)
    st.eng.fuse_gn_train = fuse
    data = synthetic_chirps(B, L, snr=12.0, seed=31)
    y = data["y_norm"]
    cond = y if cc == 1 else torch.cat([y, gaussian((B, 4, 1), seed=6).expand(B, 4, L).contiguous() * 0.3], dim=1)
    st.load_batch(data["clean_norm"].cuda(), cond.cuda(), None)
    for _ in range(steps):
        st.step(use_graph=graph)
    torch.cuda.synchronize()
    return st, {k: v.clone() for k, v in st.layout.views(st.flat_g).items()}, st.eps_hat.clone(), float(st.loss)

for B, L, in_ch, cc in [(6, 1024, 7, 5), (32, 4096, 7, 5), (32, 4096, 3, 1)]:
    s0, g0, e0, l0 = run(False, B, L, in_ch, cc)
    s1, g1, e1, l1 = run(True, B, L, in_ch, cc)
    print(f"B={B} L={L} cin={in_ch}: loss {l0:.6f} vs {l1:.6f}; eps rel {float((e0 - e1).norm() / e0.norm()):.3e}")
    tot = float(torch.cat([v.reshape(-1) for v in g0.values()]).norm())
    worst = sorted(((float((g1[k] - g0[k]).norm()) / max(float(g0[k].norm()), 0.02 * tot), k) for k in g0), reverse=True)[:6]
    for w, k in worst:
        print(f"    {w:.3e} {k}")
    ws0, ws1 = s0.eng.workspace(B, L, True), s1.eng.workspace(B, L, True)
    for li in range(7):
        print(f"    layer {li}: raw {float((ws0.raw[li].float() - ws1.raw[li].float()).norm() / ws0.raw[li].float().norm()):.2e} "
              f"out {float((ws0.out[li].float() - ws1.out[li].float()).norm() / ws0.out[li].float().norm()):.2e} "
              f"stats {float((ws0.stats[li] - ws1.stats[li]).abs().max()):.2e}")

print("--- 6 steps, eager vs graph, fused vs plain (B=32, L=4096, cin=7)")
for graph in (False, True):
    for fuse in (False, True):
        st, g, e, l = run(fuse, 32, 4096, 7, 5, steps=6, graph=graph)
        print(f"graph={graph} fuse={fuse}: loss {l:.6f} grad-norm {float(st.info[0]):.4f} |p| {float(st.flat_p.norm()):.6f}")
