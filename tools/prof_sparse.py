import os, sys
ROOT="/root/repo"
sys.path[:0]=[ROOT, ROOT+"/project-nerf_b200"]
import torch, bench
from b2n import synthetic
from src.core import NeuralField
from src.renderer import DensityGrid, render_rays
from torch.profiler import ProfilerActivity, profile
dev=torch.device("cuda",0)
torch.manual_seed(0)
model=NeuralField(bench.C2).to(dev).train()
table=model.representation.encoding.params
g=DensityGrid(resolution=128,bound=1.5,threshold=0.12).to(dev)
g.binary_grid=synthetic.ball_occupancy(128,1.5).to(dev)
opt=torch.optim.AdamW(model.parameters(),lr=0.01,weight_decay=1e-5)
bg=torch.ones(3,device=dev)
batch=tuple(t.to(dev) for t in synthetic.random_rays(2**18,seed=1))
def step():
    ro,rd,rgba=batch
    target=rgba[:,:3]*rgba[:,3:4]+bg*(1-rgba[:,3:4])
    pred,_,_=render_rays(model=model,rays_o=ro,rays_d=rd,near=2.0,far=6.0,n_samples=128,perturb=True,white_bkgd=True,density_grid=g,bg_color=bg)
    loss=torch.nn.functional.mse_loss(pred,target)+torch.mean(torch.abs(table[1:]-table[:-1]))*1e-6
    opt.zero_grad(); loss.backward()
    torch.nn.utils.clip_grad_norm_(model.representation.parameters(),1.0)
    torch.nn.utils.clip_grad_norm_(model.decoder.parameters(),1.0)
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
