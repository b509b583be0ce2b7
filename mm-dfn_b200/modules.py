"""Drop-in nn.Modules for MM-DFN's hot path, running on the sm_100a kernels of libmmdfn_b200.so.

Constructor / forward signatures and state_dict keys mirror the reference (cited per class;
paths relative to the reference root) so that code/run_train_erc.py works unchanged with
mm-dfn_b200/dropin first on sys.path.  Sub-modules are created in the reference's order with
the same nn primitives, hence the same seed yields the same initial weights.

Supported configurations (anything else raises NotImplementedError -- there is no fallback): base_model='LSTM',
av_using_lstm=False, D_e = graph_hidden_size = 100, with or without use_crn_speaker / reason_flag, and
  * multi_modal, modals='avl', graph_type 'GDF' (the MM-DFN path; att_type 'concat_subsequently' or 'mfn'), 'GF' (no
    fusion gate), 'relation' (RGCN/GraphConv per modality; att_type also 'gated'), 'DeepGCN' (GCNII per modality;
    'concat_subsequently' / 'gated' / 'mfn'), 'None' (graph-free baselines; 'concat_only' / 'lmf_only' / 'tfn_only' /
    'mfn_only' / 'gated' / 'concat_subsequently');
  * multi_modal=False (or an att_type outside the reference's multimodal list), graph_type='relation': the single-stream
    DialogueGCN configuration with the nodal-attention head."""
import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops
from .ops import DialogGeom


class BlockAdj:
    """Opaque handle of the block-compact normalised adjacency (what `create_big_adj` returns
    instead of the reference's dense (3N,3N) tensor).  `to_dense()` materialises the dense form."""

    def __init__(self, blk, diag, geom):
        self.blk, self.diag, self.geom = blk, diag, geom

    def to_dense(self):
        return ops.adj_densify(self.blk.detach(), self.diag.detach(), self.geom)

    @property
    def shape(self):
        return (3 * self.geom.N, 3 * self.geom.N)


def _side_stream(device, cache={}):
    key = str(device)
    if key not in cache:
        # high priority: the text encoder's few-CTA recurrences are the longer chain, and without it their launch waits
        # until the speaker-party encoder's 300-CTA input GEMM on the main stream has drained (measured: 55 us)
        cache[key] = torch.cuda.Stream(device=device, priority=-1)
    return cache[key]


def _geom_of(lengths, device, cache={}):
    key = (tuple(int(x) for x in lengths), str(device))
    g = cache.get(key)
    if g is None:
        if len(cache) > 64:
            cache.clear()
        g = cache[key] = DialogGeom(lengths, device)
    return g


# ------------------------------------------------------------------------------------------------
# code/model_GCN.py:157-189 (identical copy at code/model_mm.py:10-41)
# ------------------------------------------------------------------------------------------------
class GraphConvolution(nn.Module):
    def __init__(self, in_features, out_features, residual=False, variant=False):
        super().__init__()
        self.variant = variant
        self.in_features = 2 * in_features if variant else in_features
        self.out_features = out_features
        self.residual = residual
        self.weight = Parameter(torch.FloatTensor(self.in_features, self.out_features))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.out_features)
        self.weight.data.uniform_(-bound, bound)

    def forward(self, input, adj, h0, lamda, alpha, l):
        """Stand-alone layer (the model itself runs the fused GCNStackFn).  `adj` is a BlockAdj
        (message aggregate kernel) or a dense tensor (dense GEMM kernel); the two dense
        contractions run on mmdfn_gemm, the scalar mixes are elementwise glue."""
        theta = math.log(lamda / l + 1)
        if isinstance(adj, BlockAdj):
            hi = ops.SpmmFn.apply(adj.blk, adj.diag, input, adj.geom)
        else:
            hi = ops.LinearFn.apply(adj, input.t().contiguous(), None)          # adj @ input
        if self.variant:
            g = hi.shape[1]
            mm = ops.LinearFn.apply(hi, self.weight[:g].t().contiguous(), None) + \
                ops.LinearFn.apply(h0, self.weight[g:].t().contiguous(), None)
            r = (1 - alpha) * hi + alpha * h0
        else:
            r = (1 - alpha) * hi + alpha * h0
            mm = ops.LinearFn.apply(r, self.weight.t().contiguous(), None)
        out = theta * mm + (1 - theta) * r
        if self.residual:
            out = out + input
        return out


# ------------------------------------------------------------------------------------------------
# code/model_GCN.py:412-488
# ------------------------------------------------------------------------------------------------
class GCNII_lyc(nn.Module):
    def __init__(self, nfeat, nlayers, nhidden, nclass, dropout, lamda, alpha, variant, return_feature, use_residue,
                 new_graph=False, reason_flag=False):
        super().__init__()
        if (nfeat, nhidden) != (200, 100) or not variant:
            raise NotImplementedError("kernels are specialised for nfeat=200, nhidden=100, variant=True")
        self.return_feature = return_feature
        self.use_residue = use_residue
        self.new_graph = new_graph
        self.convs = nn.ModuleList([GraphConvolution(nhidden, nhidden, variant=variant) for _ in range(nlayers)])
        self.fcs = nn.ModuleList([nn.Linear(nfeat, nhidden)])
        if not return_feature:
            self.fcs.append(nn.Linear(nfeat + nhidden, nclass))
        self.act_fn = nn.ReLU()
        self.dropout = dropout
        self.alpha = alpha
        self.lamda = lamda
        self.rnn_layer = 1
        self.rnn = nn.LSTM(nhidden, nhidden, self.rnn_layer)
        self.reason_flag = reason_flag

    def forward(self, x, dia_len, topicLabel, adj=None, test_label=False, masks=None):
        """x (3N,200), adj: BlockAdj -> (3N,300) = [dropout(x) | z_K].  `masks` (tests only) injects
        keep-masks {'x','h0','layers'} (uint8) instead of drawing them."""
        if not isinstance(adj, BlockAdj):
            raise NotImplementedError("GCNII_lyc needs the block-compact adjacency from MM_GCN.create_big_adj")
        if not (self.return_feature and self.use_residue):
            raise NotImplementedError("only return_feature=True, use_residue=True (the GDF configuration)")
        geom = adj.geom
        K = len(self.convs)
        n3 = 3 * geom.N
        p = float(self.dropout)
        mx = mh = ml = None
        scale = 1.0
        if masks is not None:
            mx, mh, ml = masks.get("x"), masks.get("h0"), masks.get("layers")
            scale = 1.0 / (1.0 - p)
        elif self.training and p > 0:
            dev = x.device
            mx, mh = ops.make_mask((n3, 200), p, dev), ops.make_mask((n3, 100), p, dev)
            ml = ops.make_mask((K, n3, 100), p, dev) if K > 0 else None
            scale = 1.0 / (1.0 - p)
        with ops.sink_key("gcn"):
            return ops.GCNStackFn.apply(x, adj.blk, adj.diag, geom, K, self.reason_flag, self.lamda, self.alpha, mx, mh, ml,
                                        scale, self.fcs[0].weight, self.fcs[0].bias, self.rnn.weight_ih_l0,
                                        self.rnn.weight_hh_l0, self.rnn.bias_ih_l0, self.rnn.bias_hh_l0,
                                        *[c.weight for c in self.convs])


# ------------------------------------------------------------------------------------------------
# code/model_GCN.py:224-306
# ------------------------------------------------------------------------------------------------
class GCNII(nn.Module):
    """The per-modality deep GCN of graph_type='DeepGCN' as the reference builds it (code/model_GCN.py:225-247: same
    sub-modules in the same order).  forward(x, dia_len, qmask): x (N, 200) -> (N, 300) = [dropout(x) | dropout(z_K)] over
    the uni-modal angular-similarity graph (:274-297) -- GCNII_lyc's layer loop without the in-loop dropout, one dropout
    after it (:263-273).

    The graph kernels work on the stacked three-modality layout with ONE weight set, so a single modality runs as
    modality slot `m` of a stacked input whose cross-modal weights are zero (then the multimodal adjacency is exactly the
    three uni-modal ones): `forward_slot` takes the stacked features and the shared adjacency -- the DeepGCN model calls it
    once per modality network, paying the stack three times for the three weight sets --, `forward` stacks its input three
    times.  An ablation path: correctness first."""

    def __init__(self, nfeat, nlayers, nhidden, nclass, dropout, lamda, alpha, variant, return_feature, use_residue, new_graph=False,
                 reason_flag=False):
        super().__init__()
        if (nfeat, nhidden) != (200, 100) or not variant or new_graph or not (return_feature and use_residue):
            raise NotImplementedError("GCNII: nfeat=200, nhidden=100, variant=True, new_graph=False, return_feature=True, use_residue=True")
        self.return_feature, self.use_residue, self.new_graph = return_feature, use_residue, new_graph
        self.convs = nn.ModuleList([GraphConvolution(nhidden, nhidden, variant=variant) for _ in range(nlayers)])
        self.fcs = nn.ModuleList([nn.Linear(nfeat, nhidden)])
        self.act_fn = nn.ReLU()
        self.dropout, self.alpha, self.lamda = dropout, alpha, lamda
        self.rnn_layer = 1
        self.rnn = nn.LSTM(nhidden, nhidden, self.rnn_layer)
        self.reason_flag = reason_flag

    def forward_slot(self, X, adj, slot, masks=None):
        """X (3N, 200) stacked, adj: BlockAdj with zero cross-modal weights -> (N, 300) of modality slot `slot`.
        `masks` (tests only): {'x': (3N,200), 'h0': (3N,100), 'out': (N,100)} uint8 keep masks."""
        geom = adj.geom
        N, n3, K = geom.N, 3 * geom.N, len(self.convs)
        p = float(self.dropout)
        mx = mh = mo = None
        scale = 1.0
        if masks is not None:
            mx, mh, mo = masks.get("x"), masks.get("h0"), masks.get("out")
            scale = 1.0 / (1.0 - p)
        elif self.training and p > 0:
            mx, mh, mo = ops.make_masks([(n3, 200), (n3, 100), (N, 100)], p, X.device)
            scale = 1.0 / (1.0 - p)
        F_ = ops.GCNStackFn.apply(X, adj.blk, adj.diag, geom, K, self.reason_flag, self.lamda, self.alpha, mx, mh, None, scale,
                                  self.fcs[0].weight, self.fcs[0].bias, self.rnn.weight_ih_l0, self.rnn.weight_hh_l0,
                                  self.rnn.bias_ih_l0, self.rnn.bias_hh_l0, *[c.weight for c in self.convs])
        rows = F_[slot * N:(slot + 1) * N]
        if mo is None:
            return rows
        return torch.cat([rows[:, :200], ops.MaskScaleFn.apply(rows[:, 200:].contiguous(), mo, scale)], dim=-1)

    def forward(self, x, dia_len, qmask=None, masks=None):
        if not x.is_cuda:
            raise ops.MMDFNError("GCNII.forward needs CUDA tensors: the B200 path has no CPU fallback")
        geom = _geom_of(dia_len, x.device)
        X = torch.cat([x, x, x], dim=0)
        blk, diag = ops.AdjFn.apply(X, geom, 0.0)
        if masks is not None:
            masks = {k: (torch.cat([v, v, v], dim=0) if k in ("x", "h0") else v) for k, v in masks.items()}
        return self.forward_slot(X, BlockAdj(blk, diag, geom), 0, masks)


# ------------------------------------------------------------------------------------------------
# code/model_mm.py:44-180
# ------------------------------------------------------------------------------------------------
class MM_GCN(nn.Module):
    def __init__(self, a_dim, v_dim, l_dim, n_dim, nlayers, nhidden, nclass, dropout, lamda, alpha, variant,
                 return_feature, use_residue, new_graph='full', n_speakers=2, modals=None, use_speaker=True,
                 use_modal=False, reason_flag=False, modal_weight=1.0):
        super().__init__()
        self.return_feature = return_feature
        self.use_residue = use_residue
        self.new_graph = new_graph
        self.graph_net = GCNII_lyc(nfeat=n_dim, nlayers=nlayers, nhidden=nhidden, nclass=nclass, dropout=dropout,
                                   lamda=lamda, alpha=alpha, variant=variant, return_feature=return_feature,
                                   use_residue=use_residue, reason_flag=reason_flag)
        self.a_fc = nn.Linear(a_dim, n_dim)
        self.v_fc = nn.Linear(v_dim, n_dim)
        self.l_fc = nn.Linear(l_dim, n_dim)
        self.feature_fc = nn.Linear(n_dim * 3 + nhidden * 3, nhidden) if use_residue else nn.Linear(nhidden * 3, nhidden)
        self.final_fc = nn.Linear(nhidden, nclass)
        self.act_fn = nn.ReLU()
        self.dropout = dropout
        self.alpha = alpha
        self.lamda = lamda
        self.modals = modals
        self.modal_embeddings = nn.Embedding(3, n_dim)
        self.speaker_embeddings = nn.Embedding(n_speakers, n_dim)
        self.a_spk_embs = nn.Embedding(n_speakers, n_dim)
        self.v_spk_embs = nn.Embedding(n_speakers, n_dim)
        self.l_spk_embs = nn.Embedding(n_speakers, n_dim)
        self.use_speaker = use_speaker
        self.use_modal = use_modal
        self.modal_weight = modal_weight

    def create_big_adj(self, a, v, l, dia_len, modals, modal_weight=1.0):
        """Block-compact D^-1/2 S D^-1/2 (BlockAdj); differentiable w.r.t. a, v, l."""
        if len(modals) != 3:
            raise NotImplementedError("only modals='avl'")
        X = torch.cat([a, v, l], dim=0)
        return self.adj_of_stacked(X, _geom_of(dia_len, X.device), modal_weight)

    def adj_of_stacked(self, X, geom, modal_weight=None):
        blk, diag = ops.AdjFn.apply(X, geom, self.modal_weight if modal_weight is None else modal_weight)
        return BlockAdj(blk, diag, geom)

    def forward_stacked(self, X, geom, masks=None):
        """X (3N,200) stacked [a; v; l] -> F (3N,300)."""
        adj = self.adj_of_stacked(X, geom)
        return self.graph_net(X, None, None, adj, False, masks)

    def forward(self, a, v, l, dia_len, qmask, test_label=False):
        if self.modals is None or len(self.modals) != 3:
            raise NotImplementedError("only modals='avl'")
        if self.use_speaker or self.use_modal:
            raise NotImplementedError("use_speaker / use_modal embeddings are off on the MM-DFN path (scripts never set them)")
        X = torch.cat([a, v, l], dim=0)
        F_ = self.forward_stacked(X, _geom_of(dia_len, X.device))
        n = a.shape[0]
        out = torch.cat([F_[:n], F_[n:2 * n], F_[2 * n:3 * n]], dim=-1)
        if not self.return_feature:
            raise NotImplementedError("return_feature=False")
        return out


# ------------------------------------------------------------------------------------------------
# parameter containers for modules the GDF path constructs but never calls (state_dict parity;
# code/model.py:14-165, 420-437, 718-744).  forward() of the relation-path pieces lives in relation.py.
# ------------------------------------------------------------------------------------------------
class SimpleAttention(nn.Module):
    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = input_dim
        self.scalar = nn.Linear(input_dim, 1, bias=False)


class MatchingAttention(nn.Module):
    def __init__(self, mem_dim, cand_dim, alpha_dim=None, att_type='general'):
        super().__init__()
        self.mem_dim, self.cand_dim, self.att_type = mem_dim, cand_dim, att_type
        if att_type == 'general':
            self.transform = nn.Linear(cand_dim, mem_dim, bias=False)
        if att_type == 'general2':
            self.transform = nn.Linear(cand_dim, mem_dim, bias=True)
        elif att_type == 'concat':
            self.transform = nn.Linear(cand_dim + mem_dim, alpha_dim, bias=False)
            self.vector_prod = nn.Linear(alpha_dim, 1, bias=False)

    def forward(self, M, x, mask=None):
        """The reference calls this once per candidate position t (code/model.py:637-640).  On the B200 path the whole
        loop is one batched call on the ragged rows: `relation.attentive_node_features(..., matchatt_layer=self, ...)`."""
        raise NotImplementedError("MatchingAttention: use mmdfn_b200.relation.attentive_node_features (all candidates of all "
                                  "dialogues in one call, att_type='general2')")


class Attention(nn.Module):
    def __init__(self, embed_dim, hidden_dim=None, out_dim=None, n_head=1, score_function='dot_product', dropout=0):
        super().__init__()
        hidden_dim = embed_dim // n_head if hidden_dim is None else hidden_dim
        out_dim = embed_dim if out_dim is None else out_dim
        self.embed_dim, self.hidden_dim, self.n_head, self.score_function = embed_dim, hidden_dim, n_head, score_function
        self.w_k = nn.Linear(embed_dim, n_head * hidden_dim)
        self.w_q = nn.Linear(embed_dim, n_head * hidden_dim)
        self.proj = nn.Linear(n_head * hidden_dim, out_dim)
        self.dropout = nn.Dropout(dropout)
        if score_function == 'mlp':
            self.weight = nn.Parameter(torch.Tensor(hidden_dim * 2))
        elif score_function == 'bi_linear':
            self.weight = nn.Parameter(torch.Tensor(hidden_dim, hidden_dim))
        else:
            self.register_parameter('weight', None)
        if self.weight is not None:
            bound = 1.0 / math.sqrt(hidden_dim)
            self.weight.data.uniform_(-bound, bound)


class MaskedEdgeAttention(nn.Module):
    """code/model.py:420-471.  forward ('attn1') is provided by relation.py."""

    def __init__(self, input_dim, max_seq_len, no_cuda):
        super().__init__()
        self.input_dim = input_dim
        self.max_seq_len = max_seq_len
        self.scalar = nn.Linear(input_dim, max_seq_len, bias=False)
        self.matchatt = MatchingAttention(input_dim, input_dim, att_type='general2')
        self.simpleatt = SimpleAttention(input_dim)
        self.att = Attention(input_dim, score_function='mlp')
        self.no_cuda = no_cuda

    def forward(self, M, lengths, edge_ind):
        from .relation import masked_edge_attention
        return masked_edge_attention(self, M, lengths, edge_ind)


class MMGatedAttention(nn.Module):
    """code/model.py:718-781 (constructed by DialogueGNNModel, unused on the GDF path)."""

    def __init__(self, mem_dim, cand_dim, att_type='general'):
        super().__init__()
        self.mem_dim, self.cand_dim, self.att_type = mem_dim, cand_dim, att_type
        self.dropouta, self.dropoutv, self.dropoutl = nn.Dropout(0.5), nn.Dropout(0.5), nn.Dropout(0.5)
        if att_type == 'av_bg_fusion':
            self.transform_al = nn.Linear(mem_dim * 2, cand_dim, bias=True)
            self.scalar_al = nn.Linear(mem_dim, cand_dim)
            self.transform_vl = nn.Linear(mem_dim * 2, cand_dim, bias=True)
            self.scalar_vl = nn.Linear(mem_dim, cand_dim)
        elif att_type == 'general':
            self.transform_l = nn.Linear(mem_dim, cand_dim, bias=True)
            self.transform_v = nn.Linear(mem_dim, cand_dim, bias=True)
            self.transform_a = nn.Linear(mem_dim, cand_dim, bias=True)
            self.transform_av = nn.Linear(mem_dim * 3, 1)
            self.transform_al = nn.Linear(mem_dim * 3, 1)
            self.transform_vl = nn.Linear(mem_dim * 3, 1)

    def forward(self, a, v, l, modals=None, masks=None):
        """(N, mem_dim) x 3 -> (N, 3 * cand_dim) for att_type='general' with all three modalities (code/model.py:741-781).
        Dropout(0.5) on each input in train mode; `masks` (tests only) injects the three uint8 keep-masks (a, v, l)."""
        if self.att_type != 'general' or modals is None or not all(m in modals for m in 'avl'):
            raise NotImplementedError("MMGatedAttention: only att_type='general' with modals containing a, v and l")
        xs = [a, v, l]
        if masks is not None or self.training:
            p = 0.5
            if masks is None:
                masks = ops.make_masks([tuple(x.shape) for x in xs], p, a.device)
            xs = [ops.MaskScaleFn.apply(x, m, 1.0 / (1.0 - p)) for x, m in zip(xs, masks)]
        Pa = ops.LinearFn.apply(xs[0], self.transform_a.weight, self.transform_a.bias)
        Pv = ops.LinearFn.apply(xs[1], self.transform_v.weight, self.transform_v.bias)
        Pl = ops.LinearFn.apply(xs[2], self.transform_l.weight, self.transform_l.bias)
        # the three (1, 3D) gate weights / (1,) biases as one (3, 3D) / (3,) operand; torch.cat only moves data
        w = torch.cat([self.transform_av.weight, self.transform_al.weight, self.transform_vl.weight], dim=0)
        b = torch.cat([self.transform_av.bias, self.transform_al.bias, self.transform_vl.bias], dim=0)
        return ops.GatedFuseFn.apply(xs[0], xs[1], xs[2], Pa, Pv, Pl, w, b)


def simple_batch_graphify(features, lengths, no_cuda):
    """code/model.py:553-565: (T,B,D) -> (N,D) ragged pack; no edges on the GDF path."""
    node_features = torch.cat([features[:lengths[j], j, :] for j in range(features.size(1))], dim=0)
    return node_features, None, None, None, None


# ------------------------------------------------------------------------------------------------
# code/model.py:784-1407
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# code/model_fusion.py:123-211
# ------------------------------------------------------------------------------------------------
class TFN(nn.Module):
    """Tensor fusion network (Zadeh et al., EMNLP 2017) as the reference builds it (code/model_fusion.py:128-165): the same
    sub-modules in the same order (identical state_dict keys and seed-for-seed initial weights, incl. the 309 M-parameter
    `post_fusion_layer_1`); forward(audio_x, video_x, text_x) with three (N, 300) inputs -> (N, output_dim) runs on the
    CUDA path: sub-network Linears on mmdfn_gemm, the fusion tensor / its dropout / the first post-fusion layer in
    mmdfn_tfn_fuse_fwd / _bwd (one row chunk of the 4 MB-per-row tensor at a time), ReLU, second layer, ReLU."""

    def __init__(self, input_dims=(300, 300, 300), hidden_dims=(100, 100, 100), dropouts=0.4, post_fusion_dim=300, output_dim=300):
        super().__init__()
        if tuple(hidden_dims) != (100, 100, 100) or post_fusion_dim != 300:
            raise NotImplementedError("TFN: hidden_dims=(100, 100, 100), post_fusion_dim=300 (the reference's defaults)")
        self.audio_in, self.video_in, self.text_in = input_dims
        self.audio_hidden, self.video_hidden, self.text_hidden = hidden_dims
        self.post_fusion_dim = post_fusion_dim
        self.post_fusion_prob = dropouts
        self.audio_subnet = nn.Linear(self.audio_in, self.audio_hidden)
        self.video_subnet = nn.Linear(self.video_in, self.video_hidden)
        self.text_subnet = nn.Linear(self.text_in, self.text_hidden)
        self.post_fusion_dropout = nn.Dropout(p=self.post_fusion_prob)
        self.post_fusion_layer_1 = nn.Linear((self.text_hidden + 1) * (self.video_hidden + 1) * (self.audio_hidden + 1), self.post_fusion_dim)
        self.post_fusion_layer_2 = nn.Linear(self.post_fusion_dim, output_dim)

    def forward(self, audio_x, video_x, text_x):
        if not audio_x.is_cuda:
            raise ops.MMDFNError("TFN.forward needs CUDA tensors: the B200 path has no CPU fallback")
        ha = ops.LinearFn.apply(audio_x, self.audio_subnet.weight, self.audio_subnet.bias)
        hv = ops.LinearFn.apply(video_x, self.video_subnet.weight, self.video_subnet.bias)
        ht = ops.LinearFn.apply(text_x, self.text_subnet.weight, self.text_subnet.bias)
        p = float(self.post_fusion_dropout.p) if self.post_fusion_dropout.training else 0.0
        y1 = ops.TFNFuseFn.apply(ha, hv, ht, self.post_fusion_layer_1.weight, self.post_fusion_layer_1.bias, p)
        y1 = ops.ReluMaskFn.apply(y1, None, 1.0)
        y2 = ops.LinearFn.apply(y1, self.post_fusion_layer_2.weight, self.post_fusion_layer_2.bias)
        return ops.ReluMaskFn.apply(y2, None, 1.0)


# ------------------------------------------------------------------------------------------------
# code/model_fusion.py:214-310
# ------------------------------------------------------------------------------------------------
class LMF(nn.Module):
    """Low-rank multimodal fusion (Liu et al., ACL 2018) as the reference builds it (code/model_fusion.py:220-273): the same
    sub-modules and parameters in the same order with the same initialisers (identical state_dict keys and seed-for-seed
    initial weights); forward(audio_x, video_x, text_x) with three (N, 300) inputs -> (N, output_dim) runs on the CUDA
    path (the sub-network Linears on mmdfn_gemm, the rank products and their combination in mmdfn_lmf_fuse_fwd / _bwd).
    `post_fusion_dropout` is constructed and, as in the reference's forward, never applied."""

    def __init__(self, input_dims=(300, 300, 300), hidden_dims=(300, 300, 300), dropouts=0.4, output_dim=300, rank=4, use_softmax=False):
        super().__init__()
        if use_softmax or len(set(hidden_dims)) != 1 or rank > 8:
            raise NotImplementedError("LMF: equal hidden sizes, rank <= 8, use_softmax=False (the reference's defaults)")
        self.audio_in, self.video_in, self.text_in = input_dims
        self.audio_hidden, self.video_hidden, self.text_hidden = hidden_dims
        self.audio_subnet = nn.Linear(self.audio_in, self.audio_hidden)
        self.video_subnet = nn.Linear(self.video_in, self.video_hidden)
        self.text_subnet = nn.Linear(self.text_in, self.text_hidden)
        self.output_dim, self.rank, self.use_softmax = output_dim, rank, use_softmax
        self.audio_prob = self.video_prob = self.text_prob = self.post_fusion_prob = dropouts
        self.post_fusion_dropout = nn.Dropout(p=dropouts)
        self.audio_factor = Parameter(torch.Tensor(self.rank, self.audio_hidden + 1, self.output_dim))
        self.video_factor = Parameter(torch.Tensor(self.rank, self.video_hidden + 1, self.output_dim))
        self.text_factor = Parameter(torch.Tensor(self.rank, self.text_hidden + 1, self.output_dim))
        self.fusion_weights = Parameter(torch.Tensor(1, self.rank))
        self.fusion_bias = Parameter(torch.Tensor(1, self.output_dim))
        nn.init.xavier_normal_(self.audio_factor)
        nn.init.xavier_normal_(self.video_factor)
        nn.init.xavier_normal_(self.text_factor)
        nn.init.xavier_normal_(self.fusion_weights)
        self.fusion_bias.data.fill_(0)

    def forward(self, audio_x, video_x, text_x):
        if not audio_x.is_cuda:
            raise ops.MMDFNError("LMF.forward needs CUDA tensors: the B200 path has no CPU fallback")
        ha = ops.LinearFn.apply(audio_x, self.audio_subnet.weight, self.audio_subnet.bias)
        hv = ops.LinearFn.apply(video_x, self.video_subnet.weight, self.video_subnet.bias)
        ht = ops.LinearFn.apply(text_x, self.text_subnet.weight, self.text_subnet.bias)
        return ops.LMFFuseFn.apply(ha, hv, ht, self.audio_factor, self.video_factor, self.text_factor, self.fusion_weights,
                                   self.fusion_bias)


# ------------------------------------------------------------------------------------------------
# code/model_fusion.py:10-120
# ------------------------------------------------------------------------------------------------
class MFN(nn.Module):
    """Memory Fusion Network block (Zadeh et al., AAAI 2018) as the reference builds it (code/model_fusion.py:14-60): same
    sub-modules in the same order (identical state_dict keys and seed-for-seed initial weights); forward(x) with
    x (T, n, 3 d) -> (T, n, 400) runs on the CUDA path (mmdfn_mfn_fwd / _bwd).  `masks` (tests only): four uint8 keep
    masks (T n, 100) for the Dropout(0.2) layers after att1_fc1, att2_fc1, gamma1_fc1, gamma2_fc1."""

    def __init__(self, d=300, config=None):
        super().__init__()
        if d != 300:
            raise NotImplementedError("MFN: the CUDA path is built for d = 300 per modality (the only value the reference uses)")
        self.d_l, self.d_a, self.d_v = d, d, d
        self.dh_l, self.dh_a, self.dh_v = 100, 100, 100
        total_h_dim = self.dh_l + self.dh_a + self.dh_v
        self.mem_dim = 100
        window_dim = 2
        attInShape = total_h_dim * window_dim
        gammaInShape = attInShape + self.mem_dim
        final_out = total_h_dim + self.mem_dim
        h = 100
        self.drop_p = 0.2
        self.lstm_l = nn.LSTMCell(self.d_l, self.dh_l)
        self.lstm_a = nn.LSTMCell(self.d_a, self.dh_a)
        self.lstm_v = nn.LSTMCell(self.d_v, self.dh_v)
        self.att1_fc1 = nn.Linear(attInShape, h)
        self.att1_fc2 = nn.Linear(h, attInShape)
        self.att1_dropout = nn.Dropout(self.drop_p)
        self.att2_fc1 = nn.Linear(attInShape, h)
        self.att2_fc2 = nn.Linear(h, self.mem_dim)
        self.att2_dropout = nn.Dropout(self.drop_p)
        self.gamma1_fc1 = nn.Linear(gammaInShape, h)
        self.gamma1_fc2 = nn.Linear(h, self.mem_dim)
        self.gamma1_dropout = nn.Dropout(self.drop_p)
        self.gamma2_fc1 = nn.Linear(gammaInShape, h)
        self.gamma2_fc2 = nn.Linear(h, self.mem_dim)
        self.gamma2_dropout = nn.Dropout(self.drop_p)
        self.out_fc1 = nn.Linear(final_out, h)          # constructed, never used (as in the reference)
        self.out_fc2 = nn.Linear(h, 1)
        self.out_dropout = nn.Dropout(self.drop_p)

    def _weights(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in ops.MFN_KEYS]

    def forward(self, x, masks=None):
        if not x.is_cuda:
            raise ops.MMDFNError("MFN.forward needs CUDA tensors: the B200 path has no CPU fallback")
        T, n = x.shape[0], x.shape[1]
        scale = 1.0
        if masks is None and self.training and self.drop_p > 0:
            masks = ops.make_masks([(T * n, 100)] * 4, self.drop_p, x.device)
        if masks is not None:
            scale = 1.0 / (1.0 - self.drop_p)
        return ops.MFNFn.apply(x, masks, scale, *self._weights())


class DialogueGNNModel(nn.Module):
    def __init__(self, base_model, D_m, D_g, D_p, D_e, D_h, D_a, graph_hidden_size, n_speakers, max_seq_len,
                 window_past, window_future, n_classes=7, listener_state=False, context_attention='simple',
                 dropout_rec=0.5, dropout=0.5, nodal_attention=True, avec=False, no_cuda=False, graph_type='relation',
                 use_topic=False, alpha=0.1, lamda=0.5, multiheads=6, graph_construct='direct', use_GCN=False,
                 use_residue=True, dynamic_edge_w=False, D_m_v=512, D_m_a=100, modals='avl', att_type='gated',
                 av_using_lstm=False, Deep_GCN_nlayers=64, dataset='IEMOCAP', use_speaker=True, use_modal=False,
                 reason_flag=False, multi_modal=True, use_crn_speaker=False, speaker_weights='1-1-1', modal_weight=1.0):
        super().__init__()
        # code/model.py:822-826: fusion types outside this list switch the model to the single-stream configuration
        if att_type not in ('gated', 'concat_subsequently', 'mfn', 'mfn_only', 'tfn_only', 'lmf_only', 'concat_only'):
            multi_modal = False
        text_only = not multi_modal
        if text_only:
            # the DialogueGCN configuration (one feature stream U, `linear_` -> BiGRU `lstm` -> windowed relation graph ->
            # nodal-attention head): code/model.py:828-849, 1035-1036, 1176-1180, 1211-1212
            if base_model != 'LSTM' or graph_type != 'relation' or use_crn_speaker or use_GCN or D_e != 100 or avec:
                raise NotImplementedError("single-stream (multi_modal=False) models: base_model='LSTM', graph_type='relation', "
                                          "use_crn_speaker=False, use_GCN=False, D_e=100, avec=False")
        elif graph_type == 'None':
            # graph-free multimodal baselines (code/model.py:952-961, 1338-1405): per-modality Linear(200 -> 100) on the encoder
            # features, [that | features] per modality, one of the fusion blocks, dropout -> smax_fc -> log_softmax
            if base_model != 'LSTM' or sorted(modals) != ['a', 'l', 'v'] or av_using_lstm or D_e != 100 or graph_hidden_size != 100 \
                    or att_type not in ('concat_subsequently', 'concat_only', 'gated', 'mfn_only', 'tfn_only', 'lmf_only'):
                raise NotImplementedError("graph_type='None': base_model='LSTM', modals='avl', D_e=graph_hidden_size=100, att_type in "
                                          "concat_subsequently / concat_only / gated / mfn_only / tfn_only / lmf_only")
        elif base_model != 'LSTM' or sorted(modals) != ['a', 'l', 'v'] \
                or graph_type not in ('GDF', 'GF', 'relation', 'DeepGCN') or av_using_lstm \
                or not (att_type in ('concat_subsequently', 'mfn') or (att_type == 'gated' and graph_type in ('relation', 'DeepGCN'))) \
                or D_e != 100 or graph_hidden_size != 100 or not use_residue or (graph_type == 'relation' and use_GCN):
            raise NotImplementedError(
                "mmdfn_b200 implements the MM-DFN hot path only: base_model='LSTM', multi_modal, modals='avl', "
                "graph_type='GDF' (or 'relation' / 'None'), att_type='concat_subsequently' or 'mfn' (or 'gated' with graph_type='relation'), "
                "D_e=graph_hidden_size=100, use_residue")
        self.base_model, self.avec, self.no_cuda, self.graph_type = base_model, avec, no_cuda, graph_type
        self.alpha, self.lamda, self.multiheads, self.graph_construct = alpha, lamda, multiheads, graph_construct
        self.use_topic, self.dropout, self.use_GCN, self.use_residue = use_topic, dropout, use_GCN, use_residue
        self.dynamic_edge_w = dynamic_edge_w
        self.return_feature = True
        self.modals = [x for x in modals]
        self.use_speaker, self.use_modal = use_speaker, use_modal
        self.att_type, self.reason_flag, self.multi_modal = att_type, reason_flag, multi_modal
        self.n_speakers, self.use_crn_speaker = n_speakers, use_crn_speaker
        self.speaker_weights = list(map(float, speaker_weights.split('-')))
        self.modal_weight = modal_weight
        self.av_using_lstm = av_using_lstm
        self.use_bert_seq = False
        self.dataset = dataset
        self.window_past, self.window_future = window_past, window_future
        self.nodal_attention = nodal_attention
        if text_only:
            ms = ''.join(self.modals)
            hidden_ = 250 if len(self.modals) == 3 else 150 if ms in ('al', 'vl') else 100          # code/model.py:829-838
            self.linear_ = nn.Linear(D_m, hidden_)
            self.lstm = nn.GRU(input_size=hidden_, hidden_size=D_e, num_layers=2, bidirectional=True, dropout=dropout)
            self.rnn_parties = nn.GRU(input_size=hidden_, hidden_size=D_e, num_layers=2, bidirectional=True, dropout=dropout)
            self.att_model = MaskedEdgeAttention(2 * D_e, max_seq_len, self.no_cuda)
            from .relation import GraphNetwork
            self.graph_net = GraphNetwork(2 * D_e, n_classes, 2 * n_speakers ** 2, max_seq_len, graph_hidden_size, dropout,
                                          self.no_cuda, self.use_GCN)
            print("construct relation graph")
            self.edge_type_mapping = {}
            for j in range(n_speakers):
                for k in range(n_speakers):
                    self.edge_type_mapping[str(j) + str(k) + '0'] = len(self.edge_type_mapping)
                    self.edge_type_mapping[str(j) + str(k) + '1'] = len(self.edge_type_mapping)
            return

        self.linear_a = nn.Linear(D_m_a, 200)
        self.linear_v = nn.Linear(D_m_v, 200)
        self.linear_l = nn.Linear(D_m, 200)
        self.lstm_l = nn.GRU(input_size=200, hidden_size=D_e, num_layers=2, bidirectional=True, dropout=dropout)
        self.rnn_parties = nn.GRU(input_size=200, hidden_size=D_e, num_layers=2, bidirectional=True, dropout=dropout)
        self.att_model = MaskedEdgeAttention(2 * D_e, max_seq_len, self.no_cuda)
        if graph_type == 'relation':
            # code/model.py:917-929: one RGCN->GraphConv network per modality over the windowed speaker/temporal edges
            from .relation import GraphNetwork
            n_relations = 2 * n_speakers ** 2
            self.graph_net_a = GraphNetwork(2 * D_e, n_classes, n_relations, max_seq_len, graph_hidden_size, dropout,
                                            self.no_cuda, self.use_GCN, self.return_feature)
            self.graph_net_v = GraphNetwork(2 * D_e, n_classes, n_relations, max_seq_len, graph_hidden_size, dropout,
                                            self.no_cuda, self.use_GCN, self.return_feature)
            self.graph_net_l = GraphNetwork(2 * D_e, n_classes, n_relations, max_seq_len, graph_hidden_size, dropout,
                                            self.no_cuda, self.use_GCN, self.return_feature)
            print("construct relation graph")
        elif graph_type == 'DeepGCN':
            # code/model.py:928-940: one GCNII per modality, lamda = 0.5, alpha = 0.1 whatever the model's own values
            mk_net = lambda: GCNII(nfeat=2 * D_e, nlayers=Deep_GCN_nlayers, nhidden=graph_hidden_size, nclass=n_classes,
                                   dropout=self.dropout, lamda=0.5, alpha=0.1, variant=True, return_feature=self.return_feature,
                                   use_residue=self.use_residue, reason_flag=self.reason_flag)
            self.graph_net_a, self.graph_net_v, self.graph_net_l = mk_net(), mk_net(), mk_net()
            print("construct " + self.graph_type, "with", Deep_GCN_nlayers, "layers")
        elif graph_type == 'None':
            self.graph_net_a = nn.Linear(2 * D_e, graph_hidden_size)
            self.graph_net_v = nn.Linear(2 * D_e, graph_hidden_size)
            self.graph_net_l = nn.Linear(2 * D_e, graph_hidden_size)
            print("construct Bi-LSTM")
        else:
            self.graph_model = MM_GCN(a_dim=2 * D_e, v_dim=2 * D_e, l_dim=2 * D_e, n_dim=2 * D_e, nlayers=Deep_GCN_nlayers,
                                      nhidden=graph_hidden_size, nclass=n_classes, dropout=self.dropout, lamda=self.lamda,
                                      alpha=self.alpha, variant=True, return_feature=self.return_feature,
                                      use_residue=self.use_residue, n_speakers=n_speakers, modals=self.modals,
                                      use_speaker=self.use_speaker, use_modal=self.use_modal,
                                      reason_flag=self.reason_flag if graph_type == 'GDF' else False,      # 'GF': no fusion gate (code/model.py:944-950)
                                      modal_weight=self.modal_weight)
            print("construct " + self.graph_type)
        self.edge_type_mapping = {}
        for j in range(n_speakers):
            for k in range(n_speakers):
                self.edge_type_mapping[str(j) + str(k) + '0'] = len(self.edge_type_mapping)
                self.edge_type_mapping[str(j) + str(k) + '1'] = len(self.edge_type_mapping)
        self.gatedatt = MMGatedAttention(2 * D_e + graph_hidden_size, graph_hidden_size, att_type='general')
        self.dropout_ = nn.Dropout(self.dropout)
        # code/model.py:984-994: the gated fusion feeds 100 features per modality pair into the classifier, the memory
        # fusion network 3 x 100 hidden states + the 100-d memory
        if att_type in ('mfn', 'mfn_only'):
            self.mfn = MFN()
            self.smax_fc = nn.Linear(400, n_classes)
        elif att_type == 'tfn_only':
            self.tfn = TFN()
            self.smax_fc = nn.Linear(300, n_classes)
        elif att_type == 'lmf_only':
            self.lmf = LMF()
            self.smax_fc = nn.Linear(300, n_classes)
        elif att_type == 'concat_only':
            self.smax_fc = nn.Linear(900, n_classes)
        else:
            self.smax_fc = nn.Linear((100 if att_type == 'gated' else 300) * len(self.modals), n_classes)

    def _gru_weights(self, gru):
        return [getattr(gru, k) for k in ops.GRU_KEYS]

    def forward(self, U, qmask, umask, seq_lengths, U_a=None, U_v=None, test_label=False, masks=None):
        """U = text (T,B,D_m), U_a = audio, U_v = visual, qmask (T,B,S) -> (log_prob (N,C), None x4).
        `masks` (tests only): injected uint8 keep-masks {'gru_l','gru_p','gcn':{...},'head'}."""
        if (self.use_speaker or self.use_modal) and self.graph_type in ('GDF', 'GF'):
            raise NotImplementedError("use_speaker / use_modal are off on the MM-DFN path")
        if not U.is_cuda:
            raise ops.MMDFNError("DialogueGNNModel.forward needs CUDA tensors: the B200 path has no CPU fallback")
        if not self.multi_modal:
            return self._forward_text_only(U, qmask, umask, seq_lengths, masks)
        T, B = U.shape[0], U.shape[1]
        S = qmask.shape[2]
        dev = U.device
        geom = _geom_of(seq_lengths, dev)
        p = float(self.dropout)
        train_drop = self.training and p > 0 and masks is None
        scale = 1.0 / (1.0 - p) if (train_drop or masks is not None) else 1.0
        mk = masks or {}
        # every dropout keep-mask of the step comes from one launch (same counter stream, same bits as separate draws):
        # text GRU, party GRU, head, and -- when the graph stack uses the same rate on the GDF path -- its three masks
        nseq_p = 3 * B * S if self.use_crn_speaker else 0
        gcn = self.graph_model.graph_net if self.graph_type in ('GDF', 'GF') else None
        pool_gcn = (train_drop and gcn is not None and gcn.training and float(gcn.dropout) == p)
        pooled = None
        if train_drop:
            shapes = [(T, B, 200), (T, nseq_p, 200), (geom.N, 900)]
            if pool_gcn:
                Kc, n3 = len(gcn.convs), 3 * geom.N
                shapes += [(n3, 200), (n3, 100), (Kc, n3, 100)]
            pooled = ops.make_masks(shapes, p, dev)
        m_l = pooled[0] if train_drop else mk.get("gru_l")
        # k1: the three projections into one stacked table (a, v, l)
        with ops.sink_key("proj"):
            Utab = ops.Proj3Fn.apply(U_a, U_v, U, self.linear_a.weight, self.linear_a.bias, self.linear_v.weight,
                                     self.linear_v.bias, self.linear_l.weight, self.linear_l.bias)
        # k2: text BiGRU over the padded sequence.  It is independent of the speaker-party encoder below and both are
        # few-CTA, latency-bound recurrences, so it runs on a side stream (autograd replays the same stream in backward).
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if self.use_crn_speaker else None
        tile_l, tile_p = ops.plan_gru_tiles(T, B, 3 * B * S) if side is not None else (0, 0)
        if side is not None:
            # the text encoder needs only the text slice of the table (and the masks drawn before it): it starts as soon as
            # that projection is done, while the audio / visual projections and the speaker partition run on the main stream
            if ops.Proj3Fn.text_ready is not None:
                side.wait_event(ops.Proj3Fn.text_ready)
            else:
                side.wait_stream(main)
            with torch.cuda.stream(side), ops.gru_tile(tile_l), ops.sink_key("gru_l"):
                E_l = ops.BiGRU2Fn.apply(Utab[2].reshape(T * B, 200), None, T, B, m_l, scale, *self._gru_weights(self.lstm_l))
        else:
            with ops.sink_key("gru_l"):
                E_l = ops.BiGRU2Fn.apply(Utab[2].reshape(T * B, 200), None, T, B, m_l, scale, *self._gru_weights(self.lstm_l))
        Q = sel = pos = None
        if self.use_crn_speaker:
            # k3: shared speaker-party BiGRU over all (modality, dialogue, speaker) sequences at once
            pos, _cnt, sel, rowmap = ops.spk_partition(qmask)
            nseq = 3 * B * S
            m_p = pooled[1] if train_drop else mk.get("gru_p")
            with ops.gru_tile(tile_p), ops.sink_key("gru_p"):
                Q = ops.BiGRU2Fn.apply(Utab.reshape(3 * T * B, 200), rowmap, T, nseq, m_p, scale,
                                       *self._gru_weights(self.rnn_parties))
        if side is not None:
            main.wait_stream(side)
            E_l.record_stream(main)
        # k3/k4: scatter + speaker-weight combine + ragged pack, written as the stacked graph input
        X = ops.PartyPackFn.apply(Utab, E_l, Q, geom, sel, pos, S, tuple(self.speaker_weights))
        m_h = pooled[2] if train_drop else mk.get("head")
        if self.graph_type == 'relation':
            return self._forward_relation(X, E_l, qmask, geom, seq_lengths, umask, m_h, scale, mk.get("gated"))
        if self.graph_type == 'None':
            return self._forward_no_graph(X, geom, T, mk), None, None, None, None
        if self.graph_type == 'DeepGCN':
            return self._forward_deep_gcn(X, geom, T, m_h, scale, mk), None, None, None, None
        gm = mk.get("gcn") if masks is not None else None
        if pool_gcn:
            gm = {"x": pooled[3], "h0": pooled[4], "layers": pooled[5] if len(gcn.convs) > 0 else None}
        F_ = self.graph_model.forward_stacked(X, geom, gm)
        if self.att_type == 'mfn':
            # code/model.py:1303-1330: emotions_feat (N, 900) = [a | v | l] blocks in that order
            return self._mfn_head(F_, geom, T, (0, 1, 2), mk), None, None, None, None
        with ops.sink_key("head"):
            log_prob = ops.HeadFn.apply(F_, geom.N, m_h, scale, self.smax_fc.weight, self.smax_fc.bias)
        return log_prob, None, None, None, None

    def _forward_deep_gcn(self, X, geom, T, m_h, scale, mk):
        """graph_type='DeepGCN' (code/model.py:1244-1293): one GCNII per modality over its own uni-modal graph -> fusion (concat /
        gated / MFN) -> dropout -> ReLU -> smax_fc -> log_softmax.  The three uni-modal adjacencies are ONE call on the stacked
        features with zero cross-modal weights.  `mk` (tests only): {'gcn_a' / 'gcn_v' / 'gcn_l': GCNII masks, 'gated': {...},
        'mfn': [...], 'head': keep mask of the final dropout}."""
        N = geom.N
        blk, diag = ops.AdjFn.apply(X, geom, 0.0)
        adj = BlockAdj(blk, diag, geom)
        inject = bool(mk)
        em = [net.forward_slot(X, adj, m, mk.get("gcn_" + n) if inject else None)
              for m, (n, net) in enumerate((("a", self.graph_net_a), ("v", self.graph_net_v), ("l", self.graph_net_l)))]
        if self.att_type == 'concat_subsequently':
            with ops.sink_key("head"):
                return ops.HeadFn.apply(torch.cat(em, dim=0), N, m_h, scale, self.smax_fc.weight, self.smax_fc.bias)
        if self.att_type == 'mfn':
            return self._mfn_head(torch.cat(em, dim=0), geom, T, (2, 0, 1), mk)
        feat = self.gatedatt(em[0], em[1], em[2], self.modals, masks=(mk.get("gated") or {}).get("in"))
        p = float(self.dropout)
        m_g = mk.get("head")
        if m_g is None and self.training and p > 0 and not inject:
            m_g = ops.make_mask((N, feat.shape[1]), p, feat.device)
        feat = ops.ReluMaskFn.apply(feat, m_g, 1.0 / (1.0 - p) if m_g is not None else 1.0)
        return ops.LogSoftmaxFn.apply(ops.LinearFn.apply(feat, self.smax_fc.weight, self.smax_fc.bias))

    def _forward_no_graph(self, X, geom, T, mk):
        """graph_type='None' (code/model.py:1338-1405): emotions_m = [graph_net_m(features_m) | features_m] (N, 300) per
        modality -> fusion (concat / gated / MFN / LMF) -> dropout -> smax_fc -> log_softmax (no ReLU on this branch).
        `mk` (tests only): {'gated': {'in': ...}, 'mfn': four masks, 'head': (N, F) keep mask of the final dropout}."""
        N = geom.N
        em = []
        for m, net in enumerate((self.graph_net_a, self.graph_net_v, self.graph_net_l)):
            x = X[m * N:(m + 1) * N]
            em.append(torch.cat([ops.LinearFn.apply(x, net.weight, net.bias), x], dim=-1))
        if self.att_type in ('concat_subsequently', 'concat_only'):
            feat = torch.cat(em, dim=-1)
        elif self.att_type == 'gated':
            feat = self.gatedatt(em[0], em[1], em[2], self.modals, masks=(mk.get("gated") or {}).get("in"))
        elif self.att_type == 'lmf_only':
            feat = self.lmf(em[0], em[1], em[2])
        elif self.att_type == 'tfn_only':
            feat = self.tfn(em[0], em[1], em[2])
        else:                                                    # 'mfn_only': emotions_tmp = [l | a | v], padded per dialogue
            x = ops.MFNPackFn.apply(torch.cat(em, dim=0), geom, T, (2, 0, 1))
            feat = ops.MFNUnpadFn.apply(self.mfn(x, masks=mk.get("mfn")), geom)
        p = float(self.dropout)
        m_g = mk.get("head")
        if m_g is None and self.training and p > 0 and not mk:
            m_g = ops.make_mask((N, feat.shape[1]), p, feat.device)
        if m_g is not None:
            feat = ops.MaskScaleFn.apply(feat, m_g, 1.0 / (1.0 - p))
        return ops.LogSoftmaxFn.apply(ops.LinearFn.apply(feat, self.smax_fc.weight, self.smax_fc.bias))

    def _forward_text_only(self, U, qmask, umask, seq_lengths, masks=None):
        """Single-stream `relation` model (code/model.py:1035-1036, 1176-1180, 1211-1212): linear_ -> 2-layer BiGRU ->
        batch_graphify (windowed speaker/temporal edges, MaskedEdgeAttention norms) -> RGCN -> GraphConv -> nodal-attention
        head.  `masks` (tests only): {'gru': (T,B,200) keep mask of the inter-layer dropout, 'head': (N, hidden) keep mask}."""
        from .relation import batch_graphify
        T, B = U.shape[0], U.shape[1]
        mk = masks or {}
        p = float(self.dropout)
        x = ops.LinearFn.apply(U, self.linear_.weight, self.linear_.bias)
        m_g = mk.get("gru")
        if m_g is None and self.training and p > 0 and masks is None:
            m_g = ops.make_mask((T, B, 200), p, U.device)
        scale = 1.0 / (1.0 - p) if m_g is not None else 1.0
        emotions = ops.BiGRU2Fn.apply(x.reshape(T * B, x.shape[-1]), None, T, B, m_g, scale, *self._gru_weights(self.lstm))
        features, edge_index, edge_norm, edge_type, edge_index_lengths = batch_graphify(
            emotions, qmask, seq_lengths, self.window_past, self.window_future, self.edge_type_mapping, self.att_model, self.no_cuda)
        log_prob = self.graph_net(features, edge_index, edge_norm, edge_type, seq_lengths, umask, self.nodal_attention, self.avec,
                                  mask=mk.get("head"))
        return log_prob, edge_index, edge_norm, edge_type, edge_index_lengths

    def _mfn_head(self, F_, geom, T, perm, mk):
        """att_type='mfn' (code/model.py:1263-1291, 1303-1330): the node features, padded per dialogue to (T, B, 900), go
        through the memory fusion network; its valid rows -> dropout -> ReLU -> smax_fc (400 -> C) -> log_softmax.
        `mk` (tests only): {'mfn': four keep masks, 'mfn_head': (N, 400) keep mask}."""
        x = ops.MFNPackFn.apply(F_, geom, T, perm)
        feat = ops.MFNUnpadFn.apply(self.mfn(x, masks=mk.get("mfn")), geom)
        p = float(self.dropout)
        m_g = mk.get("mfn_head")
        if m_g is None and self.training and p > 0 and not mk:
            m_g = ops.make_mask((geom.N, 400), p, feat.device)
        feat = ops.ReluMaskFn.apply(feat, m_g, 1.0 / (1.0 - p) if m_g is not None else 1.0)
        return ops.LogSoftmaxFn.apply(ops.LinearFn.apply(feat, self.smax_fc.weight, self.smax_fc.bias))

    def _forward_relation(self, X, E_l, qmask, geom, seq_lengths, umask, m_h, scale, gated_masks=None):
        """graph_type='relation' (code/model.py:1182-1242): windowed speaker/temporal edges, edge weights from
        MaskedEdgeAttention over the padded text-branch features (the reference calls batch_graphify once per modality
        with the same att_model and keeps the last = 'l' result), one RGCNConv->GraphConv network per modality,
        concat -> dropout -> smax_fc -> log_softmax (no ReLU on this branch)."""
        from . import relation as rel
        N = geom.N
        xa, xv, xl = X[:N], X[N:2 * N], X[2 * N:]
        M_l = ops.UnpackPadFn.apply(xl, E_l, geom)                                   # padded emotions_l (T,B,200)
        edges = rel.EdgeSet(qmask, geom, self.window_past, self.window_future)
        rel._node_speakers(edges, qmask)
        edges.edge_index._mmdfn_edges = edges
        edge_norm = rel.EdgeAttnFn.apply(M_l, self.att_model.scalar.weight, edges)
        args = (edges.edge_index, edge_norm, edges.edge_type, seq_lengths, umask, self.nodal_attention, self.avec)
        if self.att_type == 'gated':
            # code/model.py:1235-1239: gatedatt over the three networks' features -> dropout -> smax_fc -> log_softmax.
            # `gated_masks` (tests only): {'in': three uint8 keep-masks of the module's Dropout(0.5), 'head': (N, 300)}
            gmk = gated_masks or {}
            feat = self.gatedatt(self.graph_net_a(xa, *args), self.graph_net_v(xv, *args), self.graph_net_l(xl, *args),
                                 self.modals, masks=gmk.get("in"))
            m_g = gmk.get("head")
            if m_g is None and self.training and float(self.dropout) > 0:
                m_g = ops.make_mask((N, feat.shape[1]), float(self.dropout), feat.device)
            if m_g is not None:
                feat = ops.MaskScaleFn.apply(feat, m_g, 1.0 / (1.0 - float(self.dropout)))
            log_prob = ops.LogSoftmaxFn.apply(ops.LinearFn.apply(feat, self.smax_fc.weight, self.smax_fc.bias))
            return log_prob, edges.edge_index, edge_norm, edges.edge_type, list(edges.counts)
        F_ = torch.cat([self.graph_net_a(xa, *args), self.graph_net_v(xv, *args), self.graph_net_l(xl, *args)], dim=0)
        if self.att_type == 'mfn':
            # code/model.py:1263-1291: emotions_tmp = [emotions_l | emotions_a | emotions_v]
            T = int(qmask.shape[0])
            return self._mfn_head(F_, geom, T, (2, 0, 1), gated_masks or {}), edges.edge_index, edge_norm, edges.edge_type, list(edges.counts)
        log_prob = ops.HeadFn.apply(F_, N, m_h, scale, self.smax_fc.weight, self.smax_fc.bias, False)
        return log_prob, edges.edge_index, edge_norm, edges.edge_type, list(edges.counts)


# ------------------------------------------------------------------------------------------------
# code/loss.py:5-34
# ------------------------------------------------------------------------------------------------
class FocalLoss(nn.Module):
    def __init__(self, gamma=0, alpha=None, size_average=True):
        super().__init__()
        self.gamma = gamma
        self.alpha = alpha
        if isinstance(alpha, (float, int)):
            self.alpha = torch.Tensor([alpha, 1 - alpha])
        if isinstance(alpha, list):
            self.alpha = torch.Tensor(alpha)
        self.size_average = size_average

    def forward(self, input, target):
        if input.dim() > 2:
            input = input.view(input.size(0), input.size(1), -1).transpose(1, 2).contiguous().view(-1, input.size(1))
        if self.alpha is not None and (self.alpha.device != input.device or self.alpha.dtype != input.dtype):
            self.alpha = self.alpha.to(device=input.device, dtype=input.dtype)
        return ops.FocalLossFn.apply(input, target, self.alpha, self.gamma, self.size_average)
