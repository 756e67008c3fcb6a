#!/bin/bash
# final tree: whole GPU suite (incl. the direct entry-point test)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2end_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2end_pytest.log
