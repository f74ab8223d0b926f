# gpurun --gpus 2 recipe: the default bench under torchrun on two GPUs of one box (weak-scaling headline + configs[2] strong block with
# utterance hashes), then the single-GPU strong block hash for comparison, and the world-size-2 reference arm behaviour
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo n2 rc=$?; tail -1 gpurun_out/bench_n2.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-vocoder --no-units --no-train-loss > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo n1 rc=$?; tail -1 gpurun_out/bench_n1.log | cut -c1-200
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
    print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strong", round(d["strong"]["value"]), d["strong"]["sha256_utt0"], d["strong"]["sha256_utt_last"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/gpu_dp_loss_check.py > gpurun_out/dp_loss_n2.log 2>&1; echo dp_loss rc=$?; grep "dp loss" gpurun_out/dp_loss_n2.log
