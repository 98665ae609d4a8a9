#!/bin/bash
# final sanity on the round-1 end-state tree
timeout 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
