#!/bin/bash
# the live-reference YOLO training test with its seeded draw, twice (must be the same case both times)
for i in 1 2; do
  timeout 100 python -m pytest tests/test_gpu_yolo.py -m gpu -q --tb=short -k "yolo_network_training_matches_live_reference" 2>&1 | tail -6 | cut -c1-300
done
