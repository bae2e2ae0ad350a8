"""``python -m gecco_b200 <gecco sub-command> ...`` — GECCO's own CLI with the CRF swapped for the B200 engine
(plug-in point: ``gecco.cli.main(crf_type=...)``, ``gecco/cli/commands/__init__.py:127-168``)."""
import sys

from .crf import main

if __name__ == "__main__":
    sys.exit(main())
