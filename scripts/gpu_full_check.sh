#!/bin/bash
# full GPU check of the working tree: parity tests, smoke, bench (train + decode + fbank + C5)
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_chk.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['decode'], d['extra']['long_c5'])"
