#!/bin/bash
# round 2, GPU call AQ: smoke() on the final code
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -5
