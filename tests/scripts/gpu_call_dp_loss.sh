mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/gpu_dp_loss_check.py > gpurun_out/dp_loss_n2.log 2>&1; echo dp_loss rc=$?; grep "dp loss" gpurun_out/dp_loss_n2.log || tail -20 gpurun_out/dp_loss_n2.log
