"""ncu --profile-from-start off ... python tools/prof_eval_trained.py [epochs]: the kernels of one Test() after `epochs` training epochs
(amazon-book shape); only the evaluation is inside the cudaProfilerStart/Stop range."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO, os.path.join(REPO, "tools")):
    sys.path.insert(0, p)
import bench_configs as bc
import utility.utility_train.batch_test as batch_test
import utility.utility_train.trainer as trainer
dev = torch.device("cuda:0")
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg, g, data, model = bc._build("LightGCN", "amazon-book", dev)
ft = model.fused_trainer(1e-3, 1024)
for _ in range(epochs):
    u, p, n = trainer.sample_epoch(data, dev)
    for s in range(0, len(u), 1024):
        ft.step(u[s:s + 1024], p[s:s + 1024], n[s:s + 1024])
for _ in range(2):
    batch_test.Test(data, model, dev, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
res = batch_test.Test(data, model, dev, cfg)
b.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("Test() ms", a.elapsed_time(b), "recall@20", res["recall"][1])
