#!/bin/bash
# timing experiments for psk_mma.cu (FB_MMA_DBG: 1 = loaders idle, 2 = MMA warps idle, 4 = MMA phase only); results are wrong by design
for d in "$@"; do
  FB_MMA_DBG=$d python bench.py --recordings 64 --steps 5 --warmup 3 --no-cpu --no-e2e --no-schemes 2>/dev/null > /tmp/mma_dbg.json
  python - "$d" <<'PY'
import json, sys
d = json.load(open("/tmp/mma_dbg.json"))
print("dbg", sys.argv[1], "kernel_ms", round(d["roofline"]["kernel_ms"], 3), "step_ms", round(d["ms_per_step"], 3), "payload_ok", d["payload_bytes_valid"] == d["payload_bytes_sent_rank0"])
PY
done
