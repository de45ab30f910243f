"""CPU oracle of the VecTree row (TEST INFRASTRUCTURE ONLY): numpy float64 restatement of
  * EuclideanCodebook.forward in training mode with importance weights (vectree/vq.py:262-306): nearest code, weighted cluster
    sizes and sums, EMA, Laplace smoothing, the direct update of `embed`;
  * VectorQuantize.forward's outputs (vq.py:379-442): straight-through value and commitment loss;
  * the k_expire replacement of vectree/vectree.py:203-205;
  * the index codec: dec2bin + np.packbits (vectree.py:119-125) and np.unpackbits + bin2dec (vectree/utils.py:33-39,106-112);
  * load_vqgaussian's reassembly (vectree/utils.py:5-65).
Pinned by tests/golden/pyref_vq.npz, produced by importing the reference's own modules on CPU (tests/golden/make_vq_golden.py)."""
import numpy as np


def sq_dists(x, embed):
    x, embed = np.asarray(x, np.float64), np.asarray(embed, np.float64)
    return (x * x).sum(1)[:, None] - 2.0 * x @ embed.T + (embed * embed).sum(1)[None, :]


def assign(x, embed):
    """(index of the nearest code, gap between the best and the second-best squared distance)"""
    d2 = sq_dists(x, embed)
    idx = d2.argmin(1)
    part = np.partition(d2, 1, axis=1)
    return idx, part[:, 1] - part[:, 0]


def ema_step(x, weight, embed, cluster_size, idx=None, decay=0.8, eps=1e-5):
    """one training forward: returns (idx, new_embed, new_cluster_size); weight None = unweighted"""
    x, embed, cluster_size = np.asarray(x, np.float64), np.asarray(embed, np.float64), np.asarray(cluster_size, np.float64)
    K = embed.shape[0]
    if idx is None:
        idx, _ = assign(x, embed)
    w = np.ones(x.shape[0]) if weight is None else np.asarray(weight, np.float64) * x.shape[0] / np.asarray(weight, np.float64).sum()
    batch = np.bincount(idx, weights=w, minlength=K)
    esum = np.zeros_like(embed)
    np.add.at(esum, idx, x * w[:, None])
    cs = cluster_size * decay + (1 - decay) * batch
    smoothed = (cs + eps) / (cs.sum() + K * eps) * cs.sum()
    new_embed = embed * decay + (1 - decay) * esum / smoothed[:, None]
    return idx, new_embed, cs


def commitment_loss(x, quantize):
    return float(((np.asarray(quantize, np.float64) - np.asarray(x, np.float64)) ** 2).mean())


def replace_least_used(embed, cluster_size, x, weight, k):
    """vectree.py:203-205: the k least-used codes become the k most important samples of the batch"""
    embed = np.array(embed, copy=True)
    least = np.argsort(cluster_size, kind="stable")[:k]
    top = np.argsort(-np.asarray(weight), kind="stable")[:k]
    embed[least] = np.asarray(x)[top]
    return embed, least, top


def pack_indices(idx, bits):
    idx = np.asarray(idx, np.int64)
    mat = ((idx[:, None] >> np.arange(bits - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
    return np.packbits(mat.reshape(-1))


def unpack_indices(packed, n, bits):
    b = np.unpackbits(np.asarray(packed, np.uint8))[:n * bits].reshape(n, bits).astype(np.int64)
    return (b << np.arange(bits - 1, -1, -1)[None, :]).sum(1)


def dequantize(files, meta):
    """load_vqgaussian (vectree/utils.py:5-65) on arrays: files = dict of the npz payloads, meta = (n, dim, K, d)"""
    n, dim, K, d = (int(v) for v in meta)
    bits = int(np.log2(K))
    non_vq = np.unpackbits(files["non_vq_mask"])[:n].astype(bool)
    vq_mask = ~non_vq
    idx = unpack_indices(files["vq_indexs"], int(vq_mask.sum()), bits)
    full = np.zeros((n, dim), np.float32)
    full[:, 0:3] = files["xyz"].astype(np.float32)
    full[:, -8:] = files["other_attribute"].astype(np.float32)
    full[vq_mask, 6:6 + d] = files["codebook"].astype(np.float32)[idx]
    full[non_vq, 6:6 + d] = files["non_vq_feats"].astype(np.float32)
    return full
