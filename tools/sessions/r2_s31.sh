#!/bin/bash
# last session: smoke + a short default bench line on the final tree
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-400
