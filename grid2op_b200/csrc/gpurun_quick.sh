set -x
nvidia-smi -L
python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -30
