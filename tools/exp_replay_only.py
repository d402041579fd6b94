import sys
sys.path[:0]=["id-grec_b200",".","tools"]
import torch, bench_configs as bc
import utility.utility_train.trainer as trainer
dev=torch.device("cuda:0"); torch.cuda.set_device(0)
cfg, g, data, model = bc._build("LightGCN", "amazon-book", dev)
B=1024
ft = model.fused_trainer(1e-3, B)
users,pos,neg = trainer.sample_epoch(data, dev)
for s in range(300): ft.step(users[s*B:(s+1)*B], pos[s*B:(s+1)*B], neg[s*B:(s+1)*B])
torch.cuda.synchronize()
def t(fn, n=1500):
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(n): fn(s)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
print("step (3 copies + replay)  ms", t(lambda s: ft.step(users[(300+s)*B:(301+s)*B], pos[(300+s)*B:(301+s)*B], neg[(300+s)*B:(301+s)*B])))
gr = ft._graphs[B]
print("replay only               ms", t(lambda s: gr.replay()))
print("step again                ms", t(lambda s: ft.step(users[(300+s)*B:(301+s)*B], pos[(300+s)*B:(301+s)*B], neg[(300+s)*B:(301+s)*B])))
print("replay only               ms", t(lambda s: gr.replay()))
