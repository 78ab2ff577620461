cd $GRAFT_REPO_ROOT
export PM_LAP_SQUARE_SLACK=1
for cfg in "4 1" "4 2" "8 1" "8 2" "2 2" "2 0" "3 3" "6 3"; do
  set -- $cfg
  echo "== eps_stop $1 last $2"
  PM_LAP_EPS_STOP_LIVE=$1 PM_LAP_EPS_STOP_LAST=$2 timeout 100 python tools/lap_slack.py 8000 0 100 2>&1 | awk '{print $2,$4,$6,$7,$8,"bids",$10,"aug",$19,"dij",$21}'
done
