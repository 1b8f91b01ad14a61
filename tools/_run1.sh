python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/dse_sweep.py --mappings 4 --multipliers 8 --steps 200 --threads 16
python tools/dse_sweep.py --mappings 16 --multipliers 8 --steps 200 --threads 16 --batched-only
