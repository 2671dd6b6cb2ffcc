#!/bin/bash
# live-reference tests with seeded draws: must pass, twice the same
for i in 1 2; do
  timeout 150 python -m pytest tests/test_gpu_generic_geometry.py tests/test_gpu_network.py tests/test_gpu_yolo.py -m gpu -q --tb=short -k "live_reference or reference_train_api" 2>&1 | tail -8 | cut -c1-300
done
