#!/bin/bash
# The reference's whole test suite against its unmodified source under oracle/refshim (CPU; ~2.5 minutes).
# usage: tools/run_reference_tests_under_standins.sh [extra pytest args]
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd /tmp && PYTEST_DISABLE_PLUGIN_AUTOLOAD=1 PYTHONPATH=$ROOT/oracle/refshim:/root/reference/src \
  python -m pytest /root/reference/tests -p no:cacheprovider -q -rf "$@"
