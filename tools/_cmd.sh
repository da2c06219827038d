timeout 600 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
