# where does the model step go? torch profiler summary of one graph-less step
import sys, torch
sys.path.insert(0, '/root/repo')
from chexpert_b200.train import TrainStep, synthetic_batch
cl = '--cl' in sys.argv
ts = TrainStep('cuda', precision='bf16', channels_last=cl, buffered='--buf' in sys.argv, fused_prologue='--fused' in sys.argv)
x, t = synthetic_batch(16, device='cuda')
if cl: x = x.contiguous(memory_format=torch.channels_last)
for _ in range(3): ts(x, t)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): ts(x, t)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90))
