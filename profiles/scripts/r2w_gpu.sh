cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_labels.py -x -q 2>&1 | tail -2
timeout 100 python -c "
import json, torch, bench
from platymatch_b200 import device as D
print(json.dumps(bench.bench_label_row(torch, D, 6554.2, False)))
" 2>&1 | tail -1
timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pm_label -s 6 -c 2 python tools/label_probe.py 2>&1 | grep -E "pm_label|duration|inst_executed|issue_active"
