import os, sys, torch
sys.path.insert(0, '.')
from morgana_b200 import workloads
from morgana_b200.fused import AcousticObjective
for B in (16, 64, 128, 256):
    ling = workloads.linguistic_batch(batch_size=B, seed=1234)
    ac = workloads.acoustic_batch(ling['n_frames'], seed=1234)
    pred, target, n = ac['pred'].cuda(), ac['target'].cuda(), ling['n_frames'].cuda()
    obj = AcousticObjective()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): obj(pred, target, n)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            obj(pred, target, n)
        for _ in range(5): g.replay()
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(50): g.replay()
        b.record(s); s.synchronize()
        t_graph = a.elapsed_time(b) / 50
        a.record(s)
        for _ in range(50): obj(pred, target, n)
        b.record(s); s.synchronize()
        t_eager = a.elapsed_time(b) / 50
    print('B=%d graph replay %.4f ms, eager python loop %.4f ms' % (B, t_graph, t_eager))
