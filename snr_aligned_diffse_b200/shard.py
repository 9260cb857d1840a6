"""Utterance sharding for multi-GPU enhancement: utterances are independent, so a list is
grouped into equal-Tpad batches first and the BATCHES are partitioned over the ranks by
longest-processing-time-first (`batch_shards`).  No collective on the data path.

Sharding utterances and batching per rank afterwards (`lpt_shards` + `bucket_batches`, the first version) hands every
rank a slice of every Tpad bucket: at 8 ranks the 824-utterance set became 14 batches per rank filled to 46 %, and
with a measured fixed cost of ~3.7 ms per batch (graph launch, small-map latency floor, host packing) against
2.0 us per utterance-frame the sweep scaled 6.5x.  Batch-level LPT keeps the batches full (7-8 per rank)."""
import numpy as np

HOP = 128


def tpad_of(length, multiple=64):
    nf = 1 + int(length) // HOP
    return multiple * ((nf + multiple - 1) // multiple)


def synthetic_lengths(n=824, seed=0, sr=16000):
    """VoiceBank-DEMAND-test-shaped lengths: lognormal(ln 2.4, 0.45) clipped to 1.5-10 s."""
    rng = np.random.default_rng(seed)
    sec = np.clip(rng.lognormal(mean=np.log(2.4), sigma=0.45, size=n), 1.5, 10.0)
    return np.round(sec * sr).astype(np.int64)


def lpt_shards(lengths, n_ranks):
    """Greedy LPT on padded frames: returns n_ranks lists of utterance indices."""
    cost = np.array([tpad_of(l) for l in lengths])
    order = np.argsort(-cost, kind="stable")
    loads = [0] * n_ranks
    shards = [[] for _ in range(n_ranks)]
    for i in order:
        r = int(np.argmin(loads))
        shards[r].append(int(i))
        loads[r] += int(cost[i])
    return shards


def bucket_batches(lengths, indices, max_batch=16, target_frames=None, cap=64):
    """Group `indices` into batches of equal Tpad: list of (tpad, [indices]).

    target_frames=None: at most `max_batch` utterances per batch.  target_frames=F: a batch holds about F padded frames --
    max(max_batch, min(cap, F // Tpad)) utterances, sizes equalised inside a bucket -- so short utterances travel in larger
    batches and the per-batch fixed cost (the ~3.7 ms latency floor of the 356-launch graph) is paid less often."""
    by = {}
    for i in indices:
        by.setdefault(tpad_of(lengths[i]), []).append(i)
    out = []
    for tpad in sorted(by, reverse=True):
        idx = by[tpad]
        mb = max_batch if not target_frames else max(max_batch, min(cap, int(target_frames) // tpad))
        n_b = -(-len(idx) // mb)
        size = -(-len(idx) // n_b)
        for k in range(0, len(idx), size):
            out.append((tpad, idx[k:k + size]))
    return out


# fixed cost of one batch in utterance-frame units: fitted on B200 from the 1-GPU and 8-GPU sweeps of the 824-utterance
# set (0.814 s for 59 batches, 0.126 s for 14 batches on the slowest rank): 3.66 ms per batch, 2.01 us per utterance-frame
BATCH_FIXED_FRAMES = 1800


def batch_cost(tpad, n_utt, fixed=BATCH_FIXED_FRAMES):
    return fixed + int(tpad) * int(n_utt)


def batch_shards(lengths, n_ranks, max_batch=16, fixed=BATCH_FIXED_FRAMES, target_frames=None):
    """Equal-Tpad batches of the whole list, assigned to ranks by greedy LPT on `batch_cost`:
    returns n_ranks lists of (tpad, [utterance indices]), each in descending-Tpad order."""
    batches = bucket_batches(lengths, range(len(lengths)), max_batch, target_frames)
    order = sorted(range(len(batches)), key=lambda k: (-batch_cost(batches[k][0], len(batches[k][1]), fixed), k))
    loads = [0] * n_ranks
    out = [[] for _ in range(n_ranks)]
    for k in order:
        r = min(range(n_ranks), key=lambda q: (loads[q], q))
        out[r].append(k)
        loads[r] += batch_cost(batches[k][0], len(batches[k][1]), fixed)
    return [[batches[k] for k in sorted(ks)] for ks in out]
