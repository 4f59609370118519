import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench
from oracle import mmssl_oracle as O
from mmssl_b200.synthetic import TripleSampler
name = sys.argv[1] if len(sys.argv) > 1 else "baby"
ds, P, feats_cpu, _, _ = bench.build_problem(name, 2022, None)
cfg = O.HotPathConfig(embed_size=ds.embed_size, n_layers=ds.n_layers, batch_size=1024)
ui, iu = O.to_torch_coo(ds.ui_norm), O.to_torch_coo(ds.iu_norm)
smp = TripleSampler(ds.train, seed=1)
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    cpu = O.CpuHotStep(P, feats_cpu[0], feats_cpu[1], (ui, iu, ui, iu, ui, iu), ds.n_items, cfg)
    ts = []
    for i in range(3):
        u, p, n = smp.sample(1024)
        t0 = time.perf_counter(); cpu.step(u, p, n); ts.append(time.perf_counter() - t0)
    print(json.dumps({"threads": nt, "s_per_step": [round(t, 3) for t in ts]}), flush=True)
