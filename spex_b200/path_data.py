"""Padded trust-path batches: the container the multi-task entry point feeds the model with.
Same constructor, attributes and methods as the reference's utility2/utils.py:3-51 (class Data):
``data = (paths, targets[, negatives])``; paths are padded with ``n_node`` (the padding user id)
to the longest path, ``mask`` marks real positions."""
from __future__ import annotations

import numpy as np


class Data:
    def __init__(self, data, n_node, shuffle=False, graph=None, test=False):
        paths = data[0]
        self.n_node = n_node
        lens = np.fromiter((len(p) for p in paths), dtype=np.int64, count=len(paths))
        self.len_max = int(lens.max()) if len(paths) else 0
        inputs = np.full((len(paths), self.len_max), n_node, dtype=np.int64)
        for r, p in enumerate(paths):
            inputs[r, : len(p)] = p
        self.inputs = inputs
        self.mask = (np.arange(self.len_max)[None, :] < lens[:, None]).astype(np.int64)
        self.targets = np.asarray(data[1])
        self.length = len(paths)
        self.shuffle = shuffle
        self.graph = graph
        self.test = test
        if test:
            self.neg = np.asarray(data[2])

    def generate_batch(self, batch_size):
        if self.shuffle:
            order = np.arange(self.length)
            np.random.shuffle(order)
            self.inputs, self.mask, self.targets = self.inputs[order], self.mask[order], self.targets[order]
            if self.test:
                self.neg = self.neg[order]
        n_batch = -(-self.length // batch_size)
        return [np.arange(b * batch_size, min((b + 1) * batch_size, self.length)) for b in range(n_batch)]

    def get_slice(self, i):
        if self.test:
            return self.inputs[i], self.mask[i], self.targets[i], self.neg[i]
        return self.inputs[i], self.mask[i], self.targets[i]
