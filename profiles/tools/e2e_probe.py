"""Which part of the end-to-end loop of bench.py costs time: optimize() alone / + upload / + download / both (C4)."""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
from test_gpu_bench_shape import build
cfg, sig, planner = build('C4', dev)
P, H, D = 512, 64, 14
means0 = planner._particle_means.clone()
h_means = means0.cpu().pin_memory()
h_trajs = [torch.empty(P, H, D).pin_memory() for _ in range(2)]
d_means = [torch.empty_like(means0) for _ in range(2)]
up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
main_stream = torch.cuda.current_stream()
ready = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]

def run(n, up, down):
    for b in range(2):
        consumed[b].record(main_stream)
        d_means[b].copy_(means0)
    def upload(i):
        b = i % 2
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(consumed[b])
            d_means[b].copy_(h_means, non_blocking=True)
            ready[b].record(up_stream)
    if up: upload(0)
    for i in range(n):
        b = i % 2
        if up:
            if i + 1 < n: upload(i + 1)
            main_stream.wait_event(ready[b])
        planner._particle_means = d_means[b]
        traj = planner.optimize(opt_iters=1)
        consumed[b].record(main_stream)
        if down:
            done = torch.cuda.Event(); done.record(main_stream)
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(done)
                h_trajs[b].copy_(traj, non_blocking=True)
            traj.record_stream(down_stream)
    torch.cuda.synchronize()

for up, down in ((False, False), (True, False), (False, True), (True, True), (False, False)):
    run(3, up, down)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); run(20, up, down); e1.record(); torch.cuda.synchronize()
    print(f'upload={up} download={down}: {e0.elapsed_time(e1) / 20:.4f} ms/step')

# duration of the 1.8 MB pinned upload itself, alone and while kernels run
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
with torch.cuda.stream(up_stream):
    a0.record(up_stream); d_means[0].copy_(h_means, non_blocking=True); a1.record(up_stream)
torch.cuda.synchronize()
print(f'upload alone: {a0.elapsed_time(a1) * 1e3:.1f} us')
planner._particle_means = d_means[1]
for _ in range(3): planner.optimize(opt_iters=1)
with torch.cuda.stream(up_stream):
    a0.record(up_stream); d_means[0].copy_(h_means, non_blocking=True); a1.record(up_stream)
torch.cuda.synchronize()
print(f'upload while three steps run: {a0.elapsed_time(a1) * 1e3:.1f} us')
b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); b0.record()
for i in range(20):
    planner.optimize(opt_iters=1)
    with torch.cuda.stream(up_stream):
        d_means[0].copy_(h_means, non_blocking=True)      # free-running uploads, no dependency on the steps
b1.record(); torch.cuda.synchronize()
print(f'20 steps with 20 independent uploads in flight: {b0.elapsed_time(b1) / 20:.4f} ms/step')

# variant: ONE copy stream, two cross-stream edges per step (main waits for the upload; the copy stream waits for the result)
side = torch.cuda.Stream()
def run2(n):
    for b in range(2):
        d_means[b].copy_(means0)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        d_means[0].copy_(h_means, non_blocking=True); ready[0].record(side)
    for i in range(n):
        b = i % 2
        main_stream.wait_event(ready[b])
        planner._particle_means = d_means[b]
        traj = planner.optimize(opt_iters=1)
        done = torch.cuda.Event(); done.record(main_stream)
        with torch.cuda.stream(side):
            # order on the copy stream: upload of problem i + 1 (its buffer was last used by step i - 1, whose download --
            # already queued here -- waited for that step), then download of result i
            if i + 1 < n:
                d_means[1 - b].copy_(h_means, non_blocking=True); ready[1 - b].record(side)
            side.wait_event(done)
            h_trajs[b].copy_(traj, non_blocking=True)
        traj.record_stream(side)
    torch.cuda.synchronize()
run2(3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); run2(20); e1.record(); torch.cuda.synchronize()
print(f'one copy stream, two edges: {e0.elapsed_time(e1) / 20:.4f} ms/step')
