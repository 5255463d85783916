"""Monte-Carlo prediction of a converted (int8) Bayes-by-backprop model, all samples of a chunk per launch.

The reference evaluates an int8 model like any other: `for i in range(samples): model(x)` (experiments/utils.py:344-355),
every forward re-drawing, re-quantising and re-packing each layer's weights on the CPU.  Here one forward of the *same
model object* carries a chunk of samples: inside `noise.sample_batch` every int8 layer draws its n weight tensors in one
sampler launch and contracts all n samples in one convolution launch; the glue between the layers (`clamp_activation`,
quantised ReLU / residual add / average pool, flatten) is elementwise or per image, so it works unchanged on activations
whose leading dimension is n*B.  Philox streams are keyed by the global sample index: the result does not depend on the
chunk size nor on how samples are sharded over GPUs, and equals a loop of single forwards under `noise.sample_index(s)`
bit for bit (for the usual case B*C*H*W % 64 == 0; see `qbn_i8_add`'s n_vec).

Same interface as `mc.MCEngine` where it matters (`predict`, `predict_sum(x, count, sample0=)`, `regression`), so
`dist.ShardedMCPredictor` shards the samples of an int8 model over GPUs with its single all-reduce."""
import torch

from . import noise, ops
from .stochastic.mcdropout.dropout import BernoulliDropout


def balanced_chunks(total, chunk):
    """Split `total` samples into ceil(total/chunk) chunks whose sizes differ by at most one."""
    if total <= 0:
        return []
    n = -(-total // max(1, chunk))
    base, extra = divmod(total, n)
    return [base + (1 if i < extra else 0) for i in range(n)]


class Int8MCEngine:
    def __init__(self, model, chunk=25, tensor_cores=True):
        layers = [m for m in model.modules() if hasattr(m, "sampled_weights")]
        if not layers:
            raise ValueError("Int8MCEngine needs a converted model (quant_utils.convert) with int8 Bayesian layers")
        if any(isinstance(m, BernoulliDropout) and m._p > 0 for m in model.modules()):
            raise NotImplementedError("sample-batched int8 MC-Dropout is not built; run the model per sample")
        self.model, self.chunk, self.regression = model, int(chunk), False
        args = getattr(model, "args", None)
        bits = int(getattr(args, "activation_precision", 8)) if args is not None else 8
        # the model clamps every activation to `bits` right after each layer (clamp_activation, src/utils.py:25-30), so
        # letting the layer clamp to the same width changes nothing — and 7-bit activations qualify for tcgen05 kind::i8
        self.act_bits = bits if tensor_cores else 8

    @torch.no_grad()
    def predict_sum(self, x, samples, sample0=0, seed=None):
        """sum over `samples` MC samples (global indices sample0..) of the model's class probabilities: [B, K] fp32."""
        if not x.is_cuda:
            raise RuntimeError("Int8MCEngine runs on CUDA tensors only (no CPU fallback)")
        if seed is not None:
            noise.manual_seed(seed)
        was_training = self.model.training
        self.model.eval()
        batch, done, total = x.shape[0], 0, None
        try:
            for n in balanced_chunks(int(samples), self.chunk):
                with noise.sample_batch(n, sample0 + done, batch, self.act_bits):
                    probs = self.model(x)                                    # [n*B, K] (or [B, K] if nothing was sampled)
                if probs.shape[0] != n * batch:
                    raise RuntimeError("model output has %d rows for %d samples x batch %d" % (probs.shape[0], n, batch))
                part = ops.mc_mean(probs.reshape(n, batch, -1))
                total = part.mul_(n) if total is None else total.add_(part, alpha=n)
                done += n
        finally:
            self.model.train(was_training)
        return total

    def predict(self, x, samples, sample0=0, seed=None):
        """p-bar = mean over MC samples (experiments/utils.py:355)."""
        return self.predict_sum(x, samples, sample0, seed) / float(samples)
