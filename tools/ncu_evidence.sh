# ncu evidence of one round: per-kernel --set full captures of every stage of the hot path on C4 plus the launch lists of
# the whole pipeline.  Run under gpurun from the repository root.  The reports stay on the box (/tmp) -- only their text
# summaries (tools/ncu_summary.py), the search kernel's report and the gzipped launch lists come back in gpurun_out/.
TAG=${1:-r02}
cd $GRAFT_REPO_ROOT
R=/tmp/ncu_$TAG
mkdir -p $R
NCU="ncu --profile-from-start off --clock-control none --import-source on --set full"
$NCU -k regex:knn_fast -c 1 -f -o $R/knn python tools/prof_stage.py knn > gpurun_out/${TAG}_ncu_knn.log 2>&1
python tools/ncu_summary.py $R/knn.ncu-rep --opcodes > gpurun_out/${TAG}_knn_ncu_summary.txt
cp $R/knn.ncu-rep gpurun_out/${TAG}_knn.ncu-rep
for st in "index 40" "density 4" "edges 4" "sort 12"; do set -- $st
  $NCU -c $2 -f -o $R/$1 python tools/prof_stage.py $1 > gpurun_out/${TAG}_ncu_$1.log 2>&1
  python tools/ncu_summary.py $R/$1.ncu-rep > gpurun_out/${TAG}_$1_ncu_summary.txt
done
$NCU -k 'regex:TagPickFirst|TagHook|DeviceSelectSweep|TagSearch|TagRank|TagLink1|TagOrdering|TagWeights|TagLeafChild|TagTourSucc' -c 60 -f -o $R/aggregate python tools/prof_stage.py aggregate > gpurun_out/${TAG}_ncu_aggregate.log 2>&1
python tools/ncu_summary.py $R/aggregate.ncu-rep > gpurun_out/${TAG}_aggregate_ncu_summary.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $R/cold.csv python tools/prof_pipeline.py C4 > /dev/null 2>&1
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $R/warm.csv python tools/prof_pipeline.py C4 > /dev/null 2>&1
gzip -c $R/cold.csv > gpurun_out/${TAG}_launches_c4_cold.csv.gz
gzip -c $R/warm.csv > gpurun_out/${TAG}_launches_c4_warm.csv.gz
ls -la gpurun_out/${TAG}_* | head -30; du -sh gpurun_out
