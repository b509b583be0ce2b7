"""Golden fixture for the drop-in integration test: the UNMODIFIED reference trainer `code/run_train_erc.py`
(/root/reference, CPU, through oracle/ref_shim.py) run for 2 epochs on a small pickle in the author's IEMOCAP format
built from the 31 IEMOCAP test dialogues (first 24 = train split, last 7 = test split), with `--dropout 0` so that the
result does not depend on a random stream.  One patch besides the shim: F.dropout / nn.Dropout return a CLONE for p = 0
(torch returns an alias there and the reference's in-place `layer_inner += q`, code/model_GCN.py:472, then breaks
autograd -- SURVEY F5d).  The printed per-epoch line (code/run_train_erc.py:628-631) is parsed and stored.

    python tests/golden/make_golden_trainer.py          (build container only)
"""
import contextlib, io, json, os, re, runpy, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

ARGS = ["--dataset", "IEMOCAP", "--Deep_GCN_nlayers", "2", "--reason_flag", "--class_weight", "--gamma", "1",
        "--speaker_weights", "3-0-1", "--dropout", "0", "--epochs", "2", "--batch-size", "8", "--lr", "0.0003", "--l2", "0.0001"]
LINE = re.compile(r"epoch: (\d+), train_loss: ([-\d.naninf]+), train_acc: ([-\d.naninf]+), train_fscore: ([-\d.naninf]+), "
                  r"valid_loss: ([-\d.naninf]+), valid_acc: ([-\d.naninf]+), valid_fscore: ([-\d.naninf]+), "
                  r"test_loss: ([-\d.naninf]+), test_acc: ([-\d.naninf]+), test_fscore: ([-\d.naninf]+)")


def parse_epochs(text):
    out = []
    for m in LINE.finditer(text):
        out.append({"epoch": int(m.group(1)), "train_loss": float(m.group(2)), "train_acc": float(m.group(3)),
                    "train_fscore": float(m.group(4)), "test_loss": float(m.group(8)), "test_acc": float(m.group(9)),
                    "test_fscore": float(m.group(10))})
    return out


def main():
    import torch
    import torch.nn.functional as F
    import ref_shim
    from helpers import write_small_iemocap_pickle
    code = ref_shim.install("/root/reference/code")
    real = F.dropout
    F.dropout = lambda x, p=0.5, training=True, inplace=False: x.clone() if p == 0 else real(x, p, training, inplace)
    torch.nn.Dropout.forward = lambda self, x: x.clone() if self.p == 0 else real(x, self.p, self.training, self.inplace)
    with tempfile.TemporaryDirectory() as tmp:
        pkl = write_small_iemocap_pickle(os.path.join(tmp, "iemocap_small.pkl"))
        buf = io.StringIO()
        old = sys.argv
        sys.argv = ["run_train_erc.py", "--no_cuda", "--data_dir", pkl] + ARGS
        try:
            with contextlib.redirect_stdout(buf):
                runpy.run_path(os.path.join(code, "run_train_erc.py"), run_name="__main__")
        finally:
            sys.argv = old
    text = buf.getvalue()
    epochs = parse_epochs(text)
    assert len(epochs) == 2, text[-2000:]
    import hashlib
    sha = hashlib.sha256(open(os.path.join(code, "run_train_erc.py"), "rb").read()).hexdigest()
    json.dump({"args": ARGS, "epochs": epochs, "run_train_erc_sha256": sha}, open(os.path.join(HERE, "run_train_erc_small.json"), "w"), indent=1)
    print(json.dumps(epochs, indent=1))


if __name__ == "__main__":
    main()
