set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2k_pytest.log 2>&1; tail -3 $O/r2k_pytest.log
timeout 420 python bench.py --stages > $O/r2k_bench.json 2> $O/r2k_bench.err; tail -c 400 $O/r2k_bench.err
timeout 120 python bench.py --steps 2 --warmup 1 --lean > $O/r2k_lean.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2k_launches.csv python bench.py --steps 2 --warmup 1 --lean > $O/r2k_ncu_l.log 2>&1
timeout 60 python tools/chi2_probe.py > $O/r2k_chi2_probe.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pm_chi2_kernel -s 2 -c 1 -o $O/r2k_chi2_shell -f python tools/chi2_probe.py > $O/r2k_ncu_c1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pm_chi2_kernel -s 5 -c 1 -o $O/r2k_chi2_dense -f python tools/chi2_probe.py > $O/r2k_ncu_c2.log 2>&1
timeout 60 python tools/label_probe.py > $O/r2k_label_probe.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pm_label_stream -s 2 -c 1 -o $O/r2k_label -f python tools/label_probe.py > $O/r2k_ncu_lab.log 2>&1
PM_LAP_SQUARE_SLACK=1 timeout 250 ncu --set full --clock-control none -k regex:"pm_ls_auction_kernel|pm_ls_bulk_kernel|pm_lap_sap_sparse" -c 14 -o $O/r2k_lap_square -f python tools/lap_slack.py 8000 0 > $O/r2k_ncu_lap.log 2>&1
for f in r2k_chi2_shell r2k_chi2_dense r2k_label r2k_lap_square; do python profiles/summarise.py kernel $O/$f.ncu-rep > $O/$f.txt 2>&1; done
ls -la $O | grep r2k
