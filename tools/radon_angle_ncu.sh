O=gpurun_out/r3g; mkdir -p $O
M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tex.sum,l1tex__data_pipe_tex_wavefronts.sum,l1tex__f_wavefronts.sum,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_tex.sum,l1tex__t_requests_pipe_tex.sum,l1tex__texin_requests.sum,lts__t_sectors_srcunit_tex.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__t_output_wavefronts_pipe_tex.sum"
for lo in 0 24 44 64; do
  hi=$((lo + 7))
  ECC_ITEM_AG_LO=$lo ECC_ITEM_AG_HI=$hi ECC_HYBRID_MODE=1 N_PROJ=16 REPS=1 INTERP=2 ncu --metrics $M --clock-control none -k regex:radon_hybrid4 -c 1 --csv --log-file $O/ag$lo.csv python tools/radon_variants.py > $O/ag$lo.log 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob("gpurun_out/r3g/ag*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    h=rows[0]; ni,vi=h.index("Metric Name"),h.index("Metric Value")
    print(f, {r[ni]:r[vi] for r in rows[1:]})
PY
