"""BASELINE configs[1] as a LOOP: "IEMOCAP full training loop, 2 GCN layers + dynamic fusion, bs=32, 1 x B200" on a synthetic
pickle in the author's format with the real IEMOCAP geometry (120 train / 31 test dialogues with the real length lists,
100/1582/342-d features).  Reports epoch throughput (utterances/s over train + test passes) of
  (a) the sync-free mirror loop (mmdfn_b200.trainer) + drop-in data path + FlatAdamTrainer,
  (b) the reference's UNCHANGED code/run_train_erc.py on the drop-in modules, with the drop-in and with the reference's
      own dataloader ("report the unchanged-script number separately"),
  (c) the unmodified reference on the host CPU (one epoch).
    python tools/trainer_bench.py [--epochs 4] [--skip-cpu]   -> one JSON line on stdout"""
import argparse, json, os, pickle, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

TEST_LENGTHS = [59, 29, 44, 28, 91, 23, 84, 54, 59, 59, 75, 64, 42, 58, 43, 49, 43, 83, 52, 58, 37, 43, 34, 44, 44, 53, 40, 68, 83, 52, 28]
ARGS = ["--dataset", "IEMOCAP", "--Deep_GCN_nlayers", "2", "--reason_flag", "--class_weight", "--gamma", "1", "--speaker_weights", "3-0-1",
        "--dropout", "0.4", "--batch-size", "32", "--lr", "0.0003", "--l2", "0.0001"]


def write_pickle(path):
    rs = np.random.RandomState(0)
    lens = bench.IEMOCAP_TRAIN_LENGTHS + TEST_LENGTHS
    vids = ["d%03d" % i for i in range(len(lens))]
    f = lambda d: {v: [row for row in rs.standard_normal((L, d)).astype(np.float32)] for v, L in zip(vids, lens)}
    tup = ({v: list(range(L)) for v, L in zip(vids, lens)}, {v: [rs.choice(["M", "F"]) for _ in range(L)] for v, L in zip(vids, lens)},
           {v: [int(x) for x in rs.randint(0, 6, L)] for v, L in zip(vids, lens)}, f(100), f(1582), f(342),
           {v: [""] * L for v, L in zip(vids, lens)}, vids[:120], vids[120:])
    pickle.dump(tup, open(path, "wb"))
    return sum(lens[:120]), sum(lens[120:])


def mirror_loop(pkl, epochs):
    import mmdfn_b200
    from mmdfn_b200.dataloader import IEMOCAPDataset
    from mmdfn_b200.dp import FlatAdamTrainer
    from mmdfn_b200.trainer import train_or_eval_graph_model
    train, test = IEMOCAPDataset(pkl, True), IEMOCAPDataset(pkl, False)
    torch.manual_seed(2021)
    model = mmdfn_b200.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, n_speakers=2, max_seq_len=200, window_past=10,
                                        window_future=10, n_classes=6, dropout=0.4, graph_type="GDF", alpha=0.2, lamda=0.5,
                                        D_m_v=342, D_m_a=1582, modals="avl", att_type="concat_subsequently", Deep_GCN_nlayers=2,
                                        use_speaker=False, reason_flag=True, use_crn_speaker=True, speaker_weights="3-0-1").cuda()
    loss_f = mmdfn_b200.FocalLoss(gamma=1.0, alpha=bench.class_weights().cuda())
    tr = FlatAdamTrainer(model, loss_f, lr=3e-4, weight_decay=1e-4)
    times = []
    for e in range(epochs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tb = [train.collate_indices(b) for b in train.length_bucketed_batches(32, shuffle=True, seed=e)]      # collate inside the timed region
        r1 = train_or_eval_graph_model(model, loss_f, tb, e, True, tr, True, "avl")
        eb = [test.collate_indices(b) for b in test.length_bucketed_batches(32)]
        r2 = train_or_eval_graph_model(model, loss_f, eb, e, False, None, True, "avl")
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return times, r1[2], r2[2]


def script_run(pkl, epochs, extra, env=None):
    script = os.path.join(ROOT, "baseline", "_ref", "code", "run_train_erc.py")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_trainer.py"), script, "--data_dir", pkl, "--epochs", str(epochs)] + ARGS + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, env=env)
    if r.returncode != 0:
        return {"error": r.stderr[-600:]}
    return [float(x) for x in re.findall(r"time: ([\d.]+) sec", r.stdout)]


def cpu_reference_epoch(pkl):
    code = "import sys; sys.path.insert(0, %r); import ref_shim, runpy; c = ref_shim.install(); sys.argv = ['run_train_erc.py', '--no_cuda', '--data_dir', %r, '--epochs', '1'] + %r; runpy.run_path(c + '/run_train_erc.py', run_name='__main__')" % (
        os.path.join(ROOT, "oracle"), pkl, ARGS)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1200, env=env)
    if r.returncode != 0:
        return {"error": r.stderr[-600:]}
    return [float(x) for x in re.findall(r"time: ([\d.]+) sec", r.stdout)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--skip-cpu", action="store_true")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        pkl = os.path.join(tmp, "iemocap_shape.pkl")
        n_train, n_test = write_pickle(pkl)
        n = n_train + n_test
        out = {"workload": "BASELINE configs[1] loop: synthetic pickle with the real IEMOCAP geometry (120 train / 31 test dialogues, %d + %d utterances, "
                           "100/1582/342-d), K=2, bs=32, dropout 0.4, Adam; one epoch = train pass (fwd+bwd+Adam) + test pass (fwd)" % (n_train, n_test),
               "utterances_per_epoch": n}
        times, l_tr, l_te = mirror_loop(pkl, a.epochs)
        out["mirror_loop"] = {"epoch_s": times, "utt_per_s_best": n / min(times), "train_loss": l_tr, "test_loss": l_te,
                              "what": "mmdfn_b200.trainer + drop-in dataloader (length buckets, pinned collate inside the timed region) + FlatAdamTrainer"}
        for tag, extra in (("unchanged_script_dropin_dataloader", []), ("unchanged_script_reference_dataloader", ["--ref-dataloader"])):
            t = script_run(pkl, a.epochs, extra)
            out[tag] = {"epoch_s": t, "utt_per_s_best": n / min(t)} if isinstance(t, list) and t else {"error": t}
        if not a.skip_cpu:
            t = cpu_reference_epoch(pkl)
            out["reference_cpu"] = {"epoch_s": t, "utt_per_s": n / t[0], "cores": os.cpu_count()} if isinstance(t, list) and t else {"error": t}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
