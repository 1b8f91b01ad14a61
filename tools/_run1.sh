python -m pytest tests -x -q -m gpu 2>&1 | tail -4
SFE_FORCE_ORDERED=1 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
