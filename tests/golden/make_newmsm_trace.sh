#!/bin/bash
# Records tests/golden/newmsm_cfg1_MSMpair_ico5_single_thread_trace.txt: the labeling and the hash of every traced mesh of every
# discrete iteration of the compiled reference (oracle/_ref/newmsm_ref_trace, --numthreads=1) on the seeded synthetic case
# bench.py's "newmsm_wall_time" leg runs (config_standard_MSMpair semantics, ico5). CPU only, ~3 - 7 minutes.
# The grid lines ("G ...", 1.9 MB of hex floats) are dropped: the comparison uses the hashes and the labels.
set -e
cd "$(dirname "$0")/../.."
python tests/newmsm_e2e.py --level 5 --config MSMpair --D 1 --skip-gpu --skip-timing-cpu --cpu-trace-out /tmp/newmsm_trace_full.txt > /dev/null
grep -v "^G " /tmp/newmsm_trace_full.txt > tests/golden/newmsm_cfg1_MSMpair_ico5_single_thread_trace.txt
