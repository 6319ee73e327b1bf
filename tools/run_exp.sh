NCU="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_kernel|decode_|rmsnorm|prologue|plan_kernel|argmax|embedding|advance|rope -s 330 -c 300 --csv"
for tag in base noshared nocoop; do
  case $tag in
    base) export -n B200_NO_PREFER_SHARED B200_GEMM_NO_COOP; unset B200_NO_PREFER_SHARED B200_GEMM_NO_COOP;;
    noshared) export B200_NO_PREFER_SHARED=1; unset B200_GEMM_NO_COOP;;
    nocoop) unset B200_NO_PREFER_SHARED; export B200_GEMM_NO_COOP=1;;
  esac
  timeout 200 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', 'ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
  timeout 300 $NCU --log-file gpurun_out/launches_exp_$tag.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --attn-iters 1 > /dev/null 2>&1
done
