#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "balanced or kmeans_session_single or tensor_path or exact_matches" 2>&1 | tail -15
