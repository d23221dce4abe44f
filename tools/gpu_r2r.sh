#!/bin/bash
# Round 2, call r: the whole GPU suite + the default bench with the pipelined e2e leg.
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/ab_line.py $OUT/${TAG}_bench.json "[default]"
python -c "
import json
j=json.load(open('$OUT/${TAG}_bench.json'))
print('   ms/step %.3f e2e %.0f host_calls %.0f' % (j['ms_per_step'], j['e2e']['value'], j['e2e_host_calls']['value']))
"
tail -c 300 $OUT/${TAG}_bench.err
