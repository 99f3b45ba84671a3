"""Launch list of ONE FiLM FIT_DECODER training step (cfg-2 shape) for `ncu --profile-from-start off --metrics
gpu__time_duration.sum`: eager step (no graph) between cudaProfilerStart/Stop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from reni_b200 import RENIAutoDecoderFiLM, RENITrainer
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = 32, 36, 128
m = RENIAutoDecoderFiLM(B, N, "SO2", 256, 5, 256, 3, 3, None, False).to(dev)
tr = RENITrainer(m, "FIT_DECODER", W, lr=1e-5, cuda_graph=False)
imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
idx = torch.arange(B, device=dev)
for _ in range(4): tr.training_step((imgs, idx))
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.training_step((imgs, idx))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
