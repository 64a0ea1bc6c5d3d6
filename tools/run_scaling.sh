#!/bin/bash
# usage: tools/run_scaling.sh N [scale]  -- default bench + BASELINE configs 3/4/5 on N GPUs of this box, JSON lines into gpurun_out/
N=${1:-1}; SC=${2:-0.1}
UNITS=$(python -c "print(int(10000 * $SC))")     # config 5: the same corpus at every N (strong scaling), so the hashes compare
if [ "$N" -gt 1 ]; then TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; else TR="python"; fi
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-schemes > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
for c in 3 4 5; do
  timeout 900 $TR bench.py --gpus $N --config $c --scale $SC --units $UNITS > gpurun_out/r02_cfg${c}_n$N.json 2> gpurun_out/r02_cfg${c}_n$N.err
done
python - <<P
import json
for f in ["r02_bench_n$N"]+[f"r02_cfg{c}_n$N" for c in (3,4,5)]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        e=d.get("e2e") or {}
        print(f, d.get("n_gpus"), round(d["value"]), round(d.get("ms_per_step"),3), e.get("value"), (e.get("h2d_probe") or {}).get("aggregate_gbs"), d.get("shard_invariance_sha256"), d.get("decision_match_min_pct"), d.get("margin_rule_violations"), [(j.get("crc_ok"), j.get("bytes")) for j in (d.get("join") or [])], d.get("units_with_reference_error"))
    except Exception as ex:
        print(f, "ERR", ex, open(f"gpurun_out/{f}.err").read()[-300:])
P
