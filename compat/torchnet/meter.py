import math

import numpy as np
import torch


class AverageValueMeter(object):
    """value() -> (mean, std) of the values added so far."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.n, self.sum, self.var = 0, 0.0, 0.0

    def add(self, value, n=1):
        self.sum += value
        self.var += value * value
        self.n += n

    def value(self):
        if self.n == 0:
            return float("nan"), float("nan")
        mean = self.sum / self.n
        if self.n == 1:
            return mean, float("inf")
        return mean, math.sqrt(max(0.0, (self.var - self.n * mean * mean) / (self.n - 1.0)))


class ClassErrorMeter(object):
    """Top-1 error (or accuracy) in percent; add(output [N,C], target [N])."""

    def __init__(self, topk=(1,), accuracy=False):
        self.topk, self.accuracy = list(topk), accuracy
        self.reset()

    def reset(self):
        self.n, self.wrong = 0, 0

    def add(self, output, target):
        if torch.is_tensor(output):
            output = output.cpu().numpy()
        if torch.is_tensor(target):
            target = target.cpu().numpy()
        pred = np.argmax(np.atleast_2d(output), 1)
        target = np.atleast_1d(target)
        self.n += int(target.shape[0])
        self.wrong += int((pred != target).sum())

    def value(self, k=-1):
        err = 100.0 * self.wrong / max(self.n, 1)
        return [100.0 - err if self.accuracy else err]
