#!/bin/bash
# ring shape sweep of k_fold_grid_async (cfg = 100*threads/32 + stages)
cd "$(dirname "$0")/.."
run() { python bench.py --skip-e2e --skip-cpu --skip-extras --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; print(round(d['ms_per_step'],3), {n:(v['ms'],v['GBps']) for n,v in k.items() if 'grid' in n})"; }
for c in 803 804 1202 1203 1602 403 404; do echo "== F1 cfg=$c"; ZB_GRID_CFG_F1=$c run; done
for c in 802 602 603 403 404; do echo "== F2 cfg=$c"; ZB_GRID_CFG_F2=$c run; done
