for wl in c2_bulk_20kx200 mid_scrna_8kx4k; do
for v in "" rankcompv3.jl_b200/variants/v_ws1.so rankcompv3.jl_b200/variants/v_ws2.so rankcompv3.jl_b200/variants/v_ws3.so; do
  if [ -n "$v" ]; then export REO_CUDA_LIB=$PWD/$v; else unset REO_CUDA_LIB; fi
  echo "== $wl variant: ${v:-default}"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-probe --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['pairs'], d['roofline']['frac'])"
done; done
