# which aten::copy_ / contiguous calls remain in the model step, by input shapes and python stack
import sys, torch
sys.path.insert(0, '/root/repo')
from chexpert_b200.train import TrainStep, synthetic_batch
ts = TrainStep('cuda', precision='bf16', buffered=True, fused_prologue=True)
x, t = synthetic_batch(16, device='cuda')
for _ in range(3): ts(x, t)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    ts(x, t)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=4) if e.key in ('aten::copy_', 'aten::_to_copy', 'aten::clone', 'aten::contiguous', 'aten::add', 'aten::mul')]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:14]:
    print(e.key, e.count, round(e.device_time_total), str(e.input_shapes)[:90])
    for fr in (e.stack or [])[:4]:
        print('     ', fr[-110:])
