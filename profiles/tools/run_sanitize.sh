# compute-sanitizer over every kernel of the pair path on small batches; summary lines into gpurun_out/TAG_sanitizer.txt
TAG=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  extra=""; [ $tool = synccheck ] && extra="--num-cuda-barriers 262144"
  timeout 900 compute-sanitizer --tool $tool $extra python profiles/tools/sanitize_small.py 600 > gpurun_out/${TAG}_san_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_san_$tool.log | tail -1)"
  grep -E "lanes stats|assembled" gpurun_out/${TAG}_san_$tool.log | tail -3
done
