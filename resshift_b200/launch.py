"""Run an UNMODIFIED reference entry script (inference_resshift.py, app.py, ...) on the B200-native hot path:

    python -m resshift_b200.launch /path/to/ResShift/inference_resshift.py -i in -o out --task realsr --scale 4

It puts ``resshift_b200/overlay`` ahead of the script's directory on ``sys.path`` so that ``sampler``,
``models.unet`` and ``models.script_util`` resolve to this package (everything else — ldm, utils, datapipe,
basicsr, the rest of ``models`` — still comes from the reference tree), then executes the script as __main__.
"""
import runpy
import sys
from pathlib import Path


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = Path(sys.argv[1]).resolve()
    overlay = Path(__file__).resolve().parent / "overlay"
    repo = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(overlay), str(script.parent), str(repo)]
    sys.argv = [str(script)] + sys.argv[2:]
    runpy.run_path(str(script), run_name="__main__")


if __name__ == "__main__":
    main()
