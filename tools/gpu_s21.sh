#!/bin/bash
# session 21b: probe before the bulk fetch of the LSTM hand-off (TSASR_DEBUG_LSTM: 4 = probe off) + parity tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_predictor_gpu.py tests/test_hardening_gpu.py -q -k "predictor" > gpurun_out/s21_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/s21_tests.log
for f in 0 4; do
  echo "== TSASR_DEBUG_LSTM=$f"
  TSASR_DEBUG_LSTM=$f timeout 120 python tools/bench_predictor.py --iters 10 2>/dev/null | grep -E "lstm_seq|fwd_us_ours|fwd_bwd_us_ours"
done | tee gpurun_out/s21b_lstm_probe.txt
