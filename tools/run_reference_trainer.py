"""Launcher: run the reference's UNCHANGED trainer (`code/run_train_erc.py`) on the B200 kernels.

    python tools/run_reference_trainer.py /path/to/MM-DFN/code/run_train_erc.py [trainer arguments ...]

Python puts a script's own directory first on sys.path, so `PYTHONPATH=.../dropin python code/run_train_erc.py` would
still import the reference's `model.py`.  This launcher puts `mm-dfn_b200/dropin` (modules `model`, `model_GCN`,
`model_mm`, `loss`, `dataloader`) ahead of the script's directory and then executes the script with runpy -- no reference
file is edited.  `--ref-dataloader` keeps the reference's own `code/dataloader.py` instead of the drop-in data path."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    args = sys.argv[2:]
    dropin = os.path.join(ROOT, "mm-dfn_b200", "dropin")
    keep_ref_loader = "--ref-dataloader" in args
    if keep_ref_loader:
        args.remove("--ref-dataloader")
        import importlib.util
        spec = importlib.util.spec_from_file_location("dataloader", os.path.join(os.path.dirname(script), "dataloader.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["dataloader"] = mod
        spec.loader.exec_module(mod)
    sys.path.insert(0, dropin)
    sys.path.insert(1, os.path.dirname(script))
    sys.argv = [script] + args
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
