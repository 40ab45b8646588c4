#!/bin/bash
# one-off verification of changes that are not switched on by default yet
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PYTHONPATH=.:tests python -c "import test_gpu_parity as T; T._pending_test_error_behaviour_matches_reference_contract(); print('pending error-behaviour test ok'); T._pending_test_decode_batch_stops_when_every_utterance_has_emitted_eos(); print('pending early-stop test ok')" 2>&1 | tail -4
SSASR_STEP_GEMM_SPLITS=2 python -m pytest tests -m gpu -x -q -k "bf16 or dual or deferred or cluster" 2>&1 | tail -2
for s in 1 2 4; do echo "splits $s"; SSASR_STEP_GEMM_SPLITS=$s python scripts/speller_time.py 2>&1 | tail -1; done
for s in 1 2; do SSASR_STEP_GEMM_SPLITS=$s python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_split$s.json 2> gpurun_out/bench_split$s.err; python -c "
import json; d=json.load(open('gpurun_out/bench_split$s.json')); print('splits $s', d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['decode']['ms'], d['extra']['decode']['utt_per_s'])"; done
