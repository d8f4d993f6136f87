#!/usr/bin/env bash
# headline bench (no CPU baseline / train block) under each environment setting given as arguments ("-" = default)
set -u
mkdir -p gpurun_out
for kv in "$@"; do
  tag=$(echo "$kv" | tr -c 'A-Za-z0-9_=' '_')
  if [ "$kv" = "-" ]; then envs=""; else envs="$kv"; fi
  env $envs timeout 900 python bench.py --no-cpu-baseline --no-train > gpurun_out/bench_$tag.json 2> gpurun_out/bench.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_$tag.json').read()); print('$kv', round(d['value'],2), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['clocks'])"; tail -n 2 gpurun_out/bench.err
  cp gpurun_out/bench_layers.json gpurun_out/bench_layers_$tag.json
done
