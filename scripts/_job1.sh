cd $GRAFT_REPO_ROOT
out=gpurun_out/r2r_stages.log; : > $out
run() { echo "== $*" >> $out; env "$@" NF=6 python scripts/flight_probe.py 2>&1 | tail -1 >> $out; }
run X=0
run RSAC_EE_STAGES=46,92,300
run RSAC_EE_STAGES=46,138,300
run RSAC_EE_STAGES=46,69,115,300
run RSAC_EE_STAGES=46,80,140,300
run RSAC_EE_STAGES=32,64,128,300
run RSAC_EE_STAGES=46,92,184
run RSAC_EE_STAGES=46,69,92,138,300
run RSAC_EE_STAGES=23,46,92,184,300
run RSAC_SELECT_SPLIT=0
run X=1
cat $out
