#!/bin/bash
# one GPU, final build: ncu --set full of one remesh of the default workload (the last remesh of tools/perf_probe.py)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/perf_probe.py many1024_1024 > gpurun_out/fin_probe.log 2>&1 &&
SDM_PROFILE_LAST=1 timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_refine|k_project|k_vertex_normals|k_orient|k_edges|k_emit' -o gpurun_out/fin_prof python tools/perf_probe.py many1024_1024 > gpurun_out/fin_ncu.log 2>&1
tail -25 gpurun_out/fin_probe.log; tail -5 gpurun_out/fin_ncu.log; ls -la gpurun_out/fin_prof.ncu-rep
exit 0
