"""Minimal driver for ncu: a few fused training steps of BASELINE configs[1] (N=36, 32 maps x 64x128), nothing else."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight
from reni_b200 import functional as F_
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = int(os.environ.get("RENI_B", "32")), 36, 128
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, mode == "latent").to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = torch.randn(B, N, 3, device=dev)
ws = F_.Workspace()
if mode == "film":  # the default FiLM decoder (5 FiLM layers, 3 x 256 mapping network): autograd training steps
    from reni_b200 import RENIAutoDecoderFiLM, RENITrainer
    mf = RENIAutoDecoderFiLM(B, N, "SO2", 256, 5, 256, 3, 3, None, False).to(dev)
    tr = RENITrainer(mf, "FIT_DECODER", W)
    imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
    for _ in range(steps):
        tr.training_step((imgs, torch.arange(B, device=dev)))
    torch.cuda.synchronize()
    print("done", mode, steps)
    sys.exit(0)
for _ in range(steps):
    if mode == "infer":
        with torch.no_grad():
            m(Z, D)
    else:
        ws.prepared_key = None
        F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(),
                                 alpha=1e-7 if mode == "latent" else 0.0, beta=1e-4 if mode == "latent" else 0.0,
                                 use_cosine=mode == "latent", need_dw=mode == "train")
torch.cuda.synchronize()
print("done", mode, steps)
