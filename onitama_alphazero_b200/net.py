"""Stand-in for the reference's policy/value network (alphazero-training/src/net.rs:14-232), kept a BLACK BOX by the
hot path: it reads ONB_BUF_LEAF_PLANES [n,21,5,5] and produces policy [n,2,25] (softmax over 50) and value [n,1] (tanh).
Harness only (plain PyTorch, library kernels): the product is the search/env path around it. Parameter names follow the
reference's VarStore paths ('resnet_0|resnet_small_block1|small_block_conv|weight', ...) so the shipped .ot archives load."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _SmallBlock(nn.Module):  # net.rs:9-38
    def __init__(self, c_in, c_out):
        super().__init__()
        self.small_block_conv = nn.Conv2d(c_in, c_out, 3, stride=1, padding=1)
        self.small_block_bn = nn.BatchNorm2d(c_out)

    def forward(self, x):
        return self.small_block_bn(self.small_block_conv(x))


class _ResNetBlock(nn.Module):  # net.rs:40-66
    def __init__(self, c):
        super().__init__()
        self.resnet_small_block1 = _SmallBlock(c, c)
        self.resnet_small_block2 = _SmallBlock(c, c)

    def forward(self, x):
        return F.relu(self.resnet_small_block2(F.relu(self.resnet_small_block1(x))) + x)


class ConvResNet(nn.Module):
    """ConvResNetConfig{hidden_channels, input_channels, resnet_block_amnt} (net.rs:74-90)."""

    def __init__(self, hidden_channels=64, input_channels=21, resnet_block_amnt=5):  # ConvResNetConfig::default (net.rs:82-90)
        super().__init__()
        h = hidden_channels
        self.conv_init_1 = nn.Conv2d(input_channels, h, 3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(h)
        for i in range(resnet_block_amnt):
            setattr(self, "resnet_%d" % i, _ResNetBlock(h))
        self.n_blocks = resnet_block_amnt
        self.vh_conv = nn.Conv2d(h, 1, 1)
        self.vh_bn = nn.BatchNorm2d(1)
        self.vh_linear1 = nn.Linear(25, h)
        self.vh_linear2 = nn.Linear(h, 1)
        self.policy_conv = nn.Conv2d(h, 2, 1)
        self.policy_bn = nn.BatchNorm2d(2)
        self.ph_linear2 = nn.Linear(50, 50)

    def forward(self, x):
        """x [n,21,5,5] -> (policy [n,2,25], value [n,1])  (net.rs:215-232; the batch dimension replaces unsqueeze(0))"""
        y = F.relu(self.bn1(self.conv_init_1(x)))
        for i in range(self.n_blocks):
            y = getattr(self, "resnet_%d" % i)(y)
        v = F.relu(self.vh_bn(self.vh_conv(y))).flatten(1)
        v = torch.tanh(self.vh_linear2(F.relu(self.vh_linear1(v))))
        p = F.relu(self.policy_bn(self.policy_conv(y))).flatten(1)
        p = torch.softmax(self.ph_linear2(p), dim=-1).reshape(-1, 2, 25)
        return p, v

    def load_ot(self, path):
        """Load a libtorch named-tensor archive saved by VarStore::save (train.rs:414-430) into THIS module. The archive must hold
        exactly this architecture: missing tensors AND tensors this module has no place for (e.g. resnet_3/resnet_4 of a 5-block
        checkpoint loaded into a 3-block module) are errors, never silently dropped. `ConvResNet.from_ot(path)` builds the module
        the archive describes."""
        src = read_ot(path)
        own = self.state_dict()
        missing = [k for k in own if k not in src and not k.endswith("num_batches_tracked")]
        unexpected = [k for k in src if k not in own]
        if missing or unexpected:
            raise KeyError("%s does not match this ConvResNet (%d blocks): missing %s, unexpected %s"
                           % (path, self.n_blocks, missing[:4], unexpected[:4]))
        for k in own:
            if k in src:
                if own[k].shape != src[k].shape:
                    raise KeyError("%s: tensor %s has shape %s, expected %s" % (path, k, tuple(src[k].shape), tuple(own[k].shape)))
                own[k].copy_(src[k])
        return self

    @classmethod
    def from_ot(cls, path):
        """The module an archive describes: hidden / input channels from conv_init_1.weight, the block count from the names."""
        src = read_ot(path)
        w = src["conv_init_1.weight"]
        blocks = 0
        while "resnet_%d.resnet_small_block1.small_block_conv.weight" % blocks in src:
            blocks += 1
        return cls(int(w.shape[0]), int(w.shape[1]), blocks).load_ot(path)

    def save_ot(self, path):
        """Write the archive VarStore::load expects (AlphaZeroMcts::from_model_file, alphazero_mcts/mod.rs:89-105; written by
        VarStore::save, train.rs:414-430): a TorchScript module whose parameters carry the VarStore paths with '|' separators;
        BatchNorm running statistics are parameters without gradient, num_batches_tracked is not part of a VarStore."""
        write_ot(self.state_dict(), path)
        return path


def read_ot(path):
    """name ('.'-separated) -> tensor of a libtorch named-tensor archive"""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        arch = torch.jit.load(path, map_location="cpu")
    out = {n.replace("|", "."): p.detach() for n, p in arch.named_parameters()}
    out.update({n.replace("|", "."): b.detach() for n, b in arch.named_buffers()})
    return out


class _VarStoreArchive(nn.Module):
    """holder whose parameter names are VarStore paths (tch joins path components with '|')"""

    def __init__(self, tensors):
        super().__init__()
        for name, t in tensors.items():
            trainable = not (name.endswith("running_mean") or name.endswith("running_var"))
            self.register_parameter(name.replace(".", "|"), nn.Parameter(t.detach().clone().float().contiguous(), requires_grad=trainable))


def write_ot(state_dict, path):
    import warnings
    tensors = {k: v for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.jit.script(_VarStoreArchive(tensors)).save(path)


def alphaloss(v, p, pi, z):
    """ConvResNet::alphaloss (net.rs:234-243): (value_loss, policy_loss) = (mean((z - v)^2), -mean_b(sum_card(log p * pi))).
    v [B,1], p [B,2,25], pi [B,2,25], z [B,1]; the reference sums over dim 1 only and then averages over the rest."""
    diff = z - v
    value_loss = (diff * diff).mean()
    policy_loss = -(p.log() * pi).sum(dim=1).mean()
    return value_loss, policy_loss


def sample_minibatch(n_samples, batch_size, generator=None):
    """train.rs:280-283 (`choose_multiple`): batch_size distinct sample indices, uniformly at random."""
    return torch.randperm(n_samples, generator=generator)[:batch_size]


def make_evaluator(model, channels_last=True):
    """Wrap a module as the `net` callable of Context.search / self_play: planes tensor -> (policy, value).
    channels_last: run the convolutions in NHWC (same f32/TF32 arithmetic, ~2.7x faster in cuDNN for these 5x5 maps:
    0.64 ms vs 1.76 ms per 4 096 positions on B200); the NCHW leaf buffer is re-laid-out by one small copy per call."""
    model.eval()
    on_gpu = next(model.parameters()).is_cuda
    if channels_last and on_gpu:
        model.to(memory_format=torch.channels_last)

    @torch.no_grad()
    def net(planes):
        if channels_last and on_gpu:
            planes = planes.contiguous(memory_format=torch.channels_last)
        p, v = model(planes)
        return p, v.reshape(-1)

    return net
