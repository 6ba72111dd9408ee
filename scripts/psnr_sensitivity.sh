for w in 1.0 0.85 0.9 1.1 1.25; do
  echo "== BIGW=$w"
  INF_WGRAD_BIGW=$w timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_2000_step and view_2000" 2>&1 | grep -E "AssertionError: \(|passed|failed" | cut -c1-200
done
