#!/bin/bash
# small-batch A/B helper (GPU box): tools/b1.sh "<env assignments>" <batch> <dtype> [steps]
env $1 python bench.py --batch $2 --dtype $3 --steps ${4:-8} --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 B=$2 $3:', round(d['ms_per_step'],2), 'ms/step', round(d['value'],3), 'clips/s')"
