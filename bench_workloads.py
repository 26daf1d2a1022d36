"""Model specs of BASELINE.json's configs, shared by bench.py (full size), the tests (reduced size) and the oracle's
model interpreter (oracle/model_ref.py).  A spec is a list of tuples:

  ("conv", cin, cout, k, padding, stride, bias)   ("linear", cin, cout, bias)
  ("bn2d", c)  ("bn1d", c)  ("relu",)  ("maxpool", k)  ("avgpool", k)  ("dropout", p)  ("flatten",)
  ("residual", [block specs], [projection specs] | None)

`build(spec)` instantiates it with compyute_b200.nn (the reference's layer classes have the same constructors, so the same
spec describes the reference model: examples/example_cnn_mnist.ipynb cell 7 for config 1; configs 3-5 are defined in
SURVEY §8d because the reference ships no VGG / ResNet / MLP)."""

from __future__ import annotations


def mnist_cnn(drop: float = 0.25):
    """Config 1: the MNIST CNN of examples/example_cnn_mnist.ipynb (cell 7); x = (B, 1, 28, 28)."""
    return [("conv", 1, 32, 5, 0, 1, True), ("relu",), ("conv", 32, 32, 5, 0, 1, False), ("bn2d", 32), ("relu",), ("maxpool", 2),
            ("dropout", drop), ("conv", 32, 64, 3, 0, 1, True), ("relu",), ("conv", 64, 64, 3, 0, 1, False), ("bn2d", 64), ("relu",),
            ("maxpool", 2), ("dropout", drop), ("flatten",), ("linear", 576, 256, False), ("bn1d", 256), ("relu",),
            ("linear", 256, 128, False), ("bn1d", 128), ("relu",), ("linear", 128, 84, False), ("bn1d", 84), ("relu",),
            ("dropout", drop), ("linear", 84, 10, True)]


def vgg(width: int = 64, in_ch: int = 3, hw: int = 32, classes: int = 10, hidden: int = 512):
    """Config 3: [Conv3x3 same - BN - ReLU] x2 at w, 2w, 4w, 8w with MaxPool2 after each stage, then 2 Linear layers."""
    spec, c = [], in_ch
    for mult in (1, 2, 4, 8):
        for _ in range(2):
            spec += [("conv", c, width * mult, 3, 1, 1, True), ("bn2d", width * mult), ("relu",)]
            c = width * mult
        spec.append(("maxpool", 2))
    feat = c * (hw // 16) * (hw // 16)
    return spec + [("flatten",), ("linear", feat, hidden, True), ("relu",), ("linear", hidden, classes, True)]


def resnet18(width: int = 64, in_ch: int = 3, hw: int = 224, classes: int = 1000):
    """Config 4: ResNet-18 shape from Compyute parts (k=2 max pool instead of 3x3/s2)."""
    def block(cin, cout, stride):
        body = [("conv", cin, cout, 3, 1, stride, False), ("bn2d", cout), ("relu",), ("conv", cout, cout, 3, 1, 1, False), ("bn2d", cout)]
        proj = [("conv", cin, cout, 1, 0, stride, False), ("bn2d", cout)] if (stride != 1 or cin != cout) else None
        return [("residual", body, proj), ("relu",)]
    spec = [("conv", in_ch, width, 7, 3, 2, False), ("bn2d", width), ("relu",), ("maxpool", 2)]
    c = width
    for i, mult in enumerate((1, 2, 4, 8)):
        spec += block(c, width * mult, 1 if i == 0 else 2) + block(width * mult, width * mult, 1)
        c = width * mult
    final = hw // 32
    return spec + [("avgpool", final), ("flatten",), ("linear", c, classes, True)]


def mlp(width: int = 4096, depth: int = 8):
    """Config 5: depth x Linear(width, width) with ReLU between; the last layer's outputs are the logits."""
    spec = []
    for i in range(depth):
        spec.append(("linear", width, width, True))
        if i != depth - 1:
            spec.append(("relu",))
    return spec


def build(spec):
    """Spec -> compyute_b200.nn.Sequential (parameters are drawn from numpy's legacy global stream, like the reference)."""
    from compyute_b200 import nn

    def mk(s):
        kind = s[0]
        if kind == "conv":
            _, cin, cout, k, pad, stride, bias = s
            return nn.Conv2D(cin, cout, k, padding=pad, stride=stride, bias=bias)
        if kind == "linear":
            return nn.Linear(s[1], s[2], bias=s[3])
        if kind == "bn2d":
            return nn.BatchNorm2D(s[1])
        if kind == "bn1d":
            return nn.BatchNorm1D(s[1])
        if kind == "relu":
            return nn.ReLU()
        if kind == "maxpool":
            return nn.MaxPooling2D(s[1])
        if kind == "avgpool":
            return nn.AvgPooling2D(s[1])
        if kind == "dropout":
            return nn.Dropout(s[1])
        if kind == "flatten":
            return nn.Flatten()
        if kind == "residual":
            proj = None
            if s[2]:
                proj = nn.Sequential(*[mk(q) for q in s[2]]) if len(s[2]) > 1 else mk(s[2][0])
            return nn.ResidualConnection(*[mk(q) for q in s[1]], residual_proj=proj)
        raise ValueError(kind)

    return nn.Sequential(*[mk(s) for s in spec])


def train_flops_per_image(spec, hw: int) -> float:
    """Algorithmic train FLOPs per image: 3 x forward contraction FLOPs (fwd + dX + dW) of Conv2D / Linear."""
    total = 0.0

    def walk(sp, h):
        nonlocal total
        for s in sp:
            if s[0] == "conv":
                _, cin, cout, k, pad, stride, _ = s
                ho = (h + 2 * pad - k) // stride + 1
                total += 2.0 * cin * cout * k * k * ho * ho
                h = ho
            elif s[0] in ("maxpool", "avgpool"):
                h //= s[1]
            elif s[0] == "linear":
                total += 2.0 * s[1] * s[2]
            elif s[0] == "residual":
                h_in = h
                h = walk(s[1], h_in)
                if s[2]:
                    walk(s[2], h_in)
        return h

    walk(spec, hw)
    return 3.0 * total
