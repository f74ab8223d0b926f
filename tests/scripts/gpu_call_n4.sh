# gpurun --gpus 4 recipe: the default bench under torchrun on four GPUs of one box (weak-scaling headline + configs[2] strong block with
# utterance hashes); context legs that only rank 0 runs are off to keep the 4x-charged box time short
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n4.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-vocoder --no-units --no-train-loss > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err; echo n4 rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n4.log") if l.startswith("{")][-1])
print("n4 value", round(d["value"]), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "strong", round(d["strong"]["value"]), d["strong"]["sha256_utt0"], d["strong"]["sha256_utt_last"], d["clocks"])
PY
tail -3 gpurun_out/bench_n4.err
