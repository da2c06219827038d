# Copies the evidence of a tools/run_round.sh run (TAG, default r2) from gpurun_out/ into profiles/ under the round's
# names and turns the `--set full` capture into the text summary (ncu reads the report here, on the CPU box).
T=${TAG:-r2}
R=${ROUND:-r2}
G=gpurun_out
cp $G/${T}_bench.json profiles/${R}_bench_1gpu.json
cp $G/${T}_bench_ref.json profiles/${R}_bench_reference_cpu.json
cp $G/${T}_bench_nchw.json profiles/${R}_bench_nchw.json
cp $G/${T}_bench_det.json profiles/${R}_bench_deterministic.json
cp $G/${T}_launches.csv profiles/${R}_launches_bench.csv
cp $G/${T}_pyramid_levels.txt profiles/${R}_pyramid_levels.txt
{ cat $G/${T}_motion_site.txt; cat $G/${T}_loss_site.txt; } > profiles/${R}_call_sites.txt
ncu -i $G/${T}_full.ncu-rep --page raw --csv > /tmp/${T}_full.csv 2>/dev/null && python tools/ncu_keys.py /tmp/${T}_full.csv > profiles/${R}_ncu_full_summary.txt
