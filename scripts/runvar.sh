for wl in c2_bulk_20kx200 mid_scrna_8kx4k c4_scrna_30kx20k; do
for v in "" rankcompv3.jl_b200/variants/v_kw4.so rankcompv3.jl_b200/variants/v_kw16.so; do
  if [ -n "$v" ]; then export REO_CUDA_LIB=$PWD/$v; else unset REO_CUDA_LIB; fi
  echo "== $wl variant: ${v:-default}"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['roofline']['rank_bits'], d['roofline']['sample_words'], d['config']['evaluations'])"
done; done
