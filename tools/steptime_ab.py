import sys
sys.path[:0]=["id-grec_b200",".","tools"]
import torch, bench_configs as bc
import utility.utility_train.trainer as trainer
dev=torch.device("cuda:0"); torch.cuda.set_device(0)
cfg, g, data, model = bc._build("LightGCN", "amazon-book", dev)
B=1024
ft = model.fused_trainer(1e-3, B)
users,pos,neg = trainer.sample_epoch(data, dev)
nb = len(users)//B
for s in range(300): ft.step(users[s*B:(s+1)*B], pos[s*B:(s+1)*B], neg[s*B:(s+1)*B])
torch.cuda.synchronize()
for rep in range(2):
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(300, 1800): ft.step(users[s*B:(s+1)*B], pos[s*B:(s+1)*B], neg[s*B:(s+1)*B])
    b.record(); torch.cuda.synchronize()
    print("LightGCN AB ms/step", a.elapsed_time(b)/1500, [float(x) for x in ft.pop_epoch_losses()])
