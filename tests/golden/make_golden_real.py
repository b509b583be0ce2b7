"""Golden fixtures at the REAL sizes of BASELINE configs 2 and 3, from the UNMODIFIED reference on CPU fp32
(imported through oracle/ref_shim.py; run in the build container only):

  c2_iemocap_test_b31_k2 / _k16 : the IEMOCAP test loader's batch 0 (code/run_train_erc.py:480-488, batch_size 32,
        shuffle=False -> all 31 test dialogues, N = 1623, T = 91), K = 2 (BASELINE configs[1]) and K = 16 (the authors'
        setting).  Only OUTPUTS are stored (eval logits, train-mode logits / loss / per-parameter gradient summaries
        with dropout = identity); the 13 MB of inputs are re-read at test time from the staged feature pickle
        (baseline/_ref/data, git-ignored, travels to the GPU box) -- the `vids` entry pins which dialogues, in order.
  c3_meld_test_b16_k4 : the first 16 MELD test dialogues in loader order, K = 4 (BASELINE configs[2]); inputs stored.

    python tests/golden/make_golden_real.py
"""
import os, sys, pickle
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import make_golden as MG  # noqa: E402


def main():
    model_mod, loss_mod, _, _ = ref_shim.reference_modules("/root/reference/code")
    import torch.nn.functional as F
    real_dropout = F.dropout
    cw_ie = torch.FloatTensor([1 / 0.086747, 1 / 0.144406, 1 / 0.227883, 1 / 0.160585, 1 / 0.127711, 1 / 0.252668])
    data = ref_shim.ref_data_dir(ROOT)

    ie = pickle.load(open(os.path.join(data, "iemocap/IEMOCAP_features.pkl"), "rb"), encoding="latin1")
    ids, spk, labels, text, audio, visual, sent, train_vid, test_vid = ie

    def ie_sample(vid):
        return (torch.FloatTensor(text[vid]), torch.FloatTensor(visual[vid]), torch.FloatTensor(audio[vid]),
                torch.FloatTensor([[1, 0] if x == "M" else [0, 1] for x in spk[vid]]),
                torch.FloatTensor([1] * len(labels[vid])), torch.LongTensor(labels[vid]))

    vids = [x for x in test_vid][:32]                       # IEMOCAPDataset.keys order (code/dataloader.py:15), batch 0
    batch = MG.collate([ie_sample(v) for v in vids])
    for K in (2, 16):
        name = f"c2_iemocap_test_b31_k{K}"
        MG.run_case(name, model_mod, loss_mod, batch, dict(S=2, C=6, dataset="IEMOCAP", spk_w="3-0-1", K=K), True,
                    class_weights=cw_ie, gamma=1.0, store_inputs=False)
        F.dropout = real_dropout
        z = dict(np.load(os.path.join(HERE, name + ".npz")))
        z["vids"] = np.array(vids)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **z)

    me = pickle.load(open(os.path.join(data, "meld/MELD_features_raw1.pkl"), "rb"), encoding="latin1")
    ids, spk, labels, text, audio, visual, sent, train_vid, test_vid, _ = me

    def me_sample(vid):
        return (torch.FloatTensor(text[vid]), torch.FloatTensor(visual[vid]), torch.FloatTensor(audio[vid]),
                torch.FloatTensor(spk[vid]), torch.FloatTensor([1] * len(labels[vid])), torch.LongTensor(labels[vid]))

    mvids = [x for x in test_vid][:16]                      # MELDDataset.keys order, batch 0 at batch_size 16
    MG.run_case("c3_meld_test_b16_k4", model_mod, loss_mod, MG.collate([me_sample(v) for v in mvids]),
                dict(S=9, C=7, K=4, dataset="MELD", spk_w="0.5-0.5-1.5"), True, class_weights=None, gamma=1.0)
    F.dropout = real_dropout
    print("done")


if __name__ == "__main__":
    main()
