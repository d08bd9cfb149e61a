#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -x -q > $out/r2u_peer_tests.log 2>&1; echo "peer tests rc=$?"; tail -15 $out/r2u_peer_tests.log
timeout 600 python scripts/vecchia_prof.py > $out/r2u_vecchia_prof.log 2>&1; echo "vecchia prof rc=$?"
