import torch, sys
sys.path.insert(0, "/root/repo")
from diffusion_models_for_gravitational_waveform_reconstruction_b200 import whitening as W
B, L = 2048, 4096
y = torch.randn(B, L, device="cuda") * 3e-3
x = torch.randn(B, L, device="cuda") * 1e-3
for _ in range(3):
    W.whiten_train_like(y, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    W.whiten_train_like(y, x)
e1.record(); torch.cuda.synchronize()
print("whiten_train_like ms", e0.elapsed_time(e1) / 5)
