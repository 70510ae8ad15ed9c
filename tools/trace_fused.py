import os, sys, contextlib, io
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rlgym_ppo_b200.ppo import ExperienceBuffer, PPOLearner
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    lr = PPOLearner(89, 90, 0, (256,)*3, (256,)*3, (0.1,1.0), 50000, 1, 3e-4, 3e-4, 0.2, 0.01, 50000, "cuda:0")
lr.use_cuda_graph=False
rng=np.random.RandomState(0); n=50000
b=ExperienceBuffer(50000,1,"cuda:0")
b.submit_experience(rng.randn(n,89).astype(np.float32), rng.randint(0,90,n).astype(np.float32), np.full(n,-4.5,np.float32), np.zeros(n,np.float32), np.zeros((n,89),np.float32), np.zeros(n,np.float32), np.zeros(n), rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32))
lr.learn(b); lr.learn(b)
os.environ["RLPPO_FUSED_TRACE"]="1"
lr.learn(b)
