#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_probe.py > gpurun_out/r2u_e2e_probe.json 2> gpurun_out/r2u_e2e_probe.err; tail -c 300 gpurun_out/r2u_e2e_probe.err; cat gpurun_out/r2u_e2e_probe.json
