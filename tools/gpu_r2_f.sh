#!/bin/bash
# round 2, pass F: GPU test-suite (hidden-layer precisions, epoch-level training golden), then the sanitizer passes
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
bash tools/gpu_sanitize.sh r02 2>&1 | tail -20
