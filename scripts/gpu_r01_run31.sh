# r31: the adapter-level identity test and the bench line with its next_rows leg, on the committed build
set -x
mkdir -p gpurun_out
python -m pytest tests/test_cpp_adapter.py tests/test_mash.py -m gpu -x -q > gpurun_out/r31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r31_pytest.log
tail -4 gpurun_out/r31_pytest.log
python bench.py > gpurun_out/r31_bench.json 2> gpurun_out/r31_bench.err
cat gpurun_out/r31_bench.json; tail -3 gpurun_out/r31_bench.err
