# 8-GPU XL exchange experiments (one gpurun --gpus 8 call): unicast vs multicast, fused vs chunked
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29777"
A="tools/bench_xl.py --eval-users 4096 --no-one-gpu-ref"
IDG_MULTICAST=0 IDG_DIST_EXCHANGE=fused timeout 200 $TR $A > gpurun_out/xl_r2c_uni_fused.json 2> gpurun_out/xl_r2c_uni_fused.err
IDG_MULTICAST=0 IDG_DIST_EXCHANGE=chunked IDG_PUSH_CTAS=32 timeout 200 $TR $A > gpurun_out/xl_r2c_uni_chunked.json 2> gpurun_out/xl_r2c_uni_chunked.err
IDG_DIST_EXCHANGE=chunked IDG_PUSH_CTAS=64 timeout 200 $TR $A > gpurun_out/xl_r2c_mc_chunked64.json 2> gpurun_out/xl_r2c_mc_chunked64.err
grep -h -o '"ms_per_train_step": [0-9.]*' gpurun_out/xl_r2c_*.json
