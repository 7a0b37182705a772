#!/bin/bash
# pass AN: full GPU suite + smoke on the final commit of the round
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -q -m gpu > gpurun_out/an_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/an_pytest.log
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
