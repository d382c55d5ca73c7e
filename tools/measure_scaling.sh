#!/usr/bin/env bash
# One multi-GPU measurement pass (run under `gpurun --gpus N -- bash tools/measure_scaling.sh N`): the headline bench
# with both exchanges, the configs[4] sweep and the proof-of-storage commit, all into gpurun_out/.
# Column hashing (encode stores into peer column blocks) is the default path; row hashing (chunk chaining values
# re-sharded instead of the encoded matrix) is the candidate to replace it from 4 GPUs up.
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p "$OUT"
run() {  # run <tag> <script> [args...]
    local tag=$1; shift
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
        "$@" > "$OUT/$tag.json" 2> "$OUT/$tag.err" || echo "$tag failed (see $OUT/$tag.err)"
}
for mode in columns rows rows-fused; do
    run "bench_n${N}_${mode}" bench.py --gpus "$N" --steps 50 --warmup 5 --hashing "$mode"
done
for mode in columns rows; do
    run "sweep_n${N}_${mode}" tools/bench_sweep.py --steps 5 --hashing "$mode"
    run "pos_n${N}_${mode}" tools/bench_pos.py --gib 4 --steps 5 --hashing "$mode"
done
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
for mode in ("columns", "rows", "rows-fused"):
    try:
        d = json.loads(open(f"gpurun_out/bench_n{n}_{mode}.json").read().strip().splitlines()[-1])
        print(mode, "ms/commit %.3f" % d["ms_per_step"], "root ok" if d.get("root_matches_golden") else "ROOT?",
              {k: round(v, 3) for k, v in d["roofline"]["kernels_ms_per_step"].items()})
    except Exception as e:  # a failed leg leaves no JSON line
        print(mode, "no result:", e)
PY
