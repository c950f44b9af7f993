#!/bin/bash
for v in 0 1; do TRIS_DEBUG_SKIP_FINALIZE=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-190; done
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 600 -k "embedding or text_tower" 2>&1 | tail -3
