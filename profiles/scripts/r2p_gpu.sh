cd $GRAFT_REPO_ROOT
timeout 120 python -m pytest tests/test_labels.py -x -q 2>&1 | tail -2
export PM_LAP_SQUARE_SLACK=1
for d in 6 3 2 1 0; do
  echo "== dummy_phases $d"
  PM_LAP_EPS_DUMMY_PHASES=$d timeout 100 python tools/lap_slack.py 8000 200 400 2>&1 | awk '{print $2,$4,$6,$7,$8,"bids",$10,"aug",$19,"dij",$21}'
done
