P='import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; print(sys.argv[1], "%.3f G/s step %.3f ms k_step %.3f apply %.3f e2e %.3f ms" % (d["value"]/1e9, d["ms_per_step"], r["kernel_ms_per_launch"], r["apply_kernel_ms_per_launch"], d["e2e"]["ms_per_step"]))'
for i in 1 2; do
  (cd _ab_old && timeout 300 python bench.py --no-cpu-baseline --topk-users 0 > ../gpurun_out/ab_old_$i.json 2>/dev/null); python -c "$P" gpurun_out/ab_old_$i.json
  timeout 300 python bench.py --no-cpu-baseline --topk-users 0 > gpurun_out/ab_new_$i.json 2>/dev/null; python -c "$P" gpurun_out/ab_new_$i.json
done
