cd $GRAFT_REPO_ROOT
for sl in 4 1 0; do
  echo "== PM_LAP_STOP_LIVE=$sl"
  PM_LAP_STOP_LIVE=$sl timeout 80 python tools/lap_probe.py 20000 2>&1 | grep -E "^batch|dijkstra" | awk '/^batch/{print $0} /dijkstra/{print "   ", $2,$3,$4,$5,$6,$7,$8,$9,$12,$13}' | head -14
done
