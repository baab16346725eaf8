#!/bin/bash
# strong scaling of the C4 job (1e9 rays per particle in total) on 1, 2, 4, 8 GPUs of one box, and the
# multi-device tests; results under gpurun_out/.   usage: scripts/scaling_run.sh TAG
tag=${1:-r2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${tag}_multi_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_multi_pytest.txt
python bench.py --scaling strong --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_strong_n1.json 2> gpurun_out/${tag}_strong_n1.err || tail -5 gpurun_out/${tag}_strong_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --scaling strong --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_strong_n$n.json 2> gpurun_out/${tag}_strong_n$n.err || tail -5 gpurun_out/${tag}_strong_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_weak_n8.json 2> gpurun_out/${tag}_weak_n8.err || tail -5 gpurun_out/${tag}_weak_n8.err
for f in gpurun_out/${tag}_strong_n*.json gpurun_out/${tag}_weak_n8.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "n_gpus", d["n_gpus"], d["scaling"], "value %.4f G rays/s"%(d["value"]/1e9), "e2e %.4f"%(d["e2e"]["value"]/1e9), "ms/step %.1f"%d["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
