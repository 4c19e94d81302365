"""Running mean of a scalar -- same interface as the reference's lib/AverageMeter.py:1-21
(attributes ``val``, ``avg``, ``sum``, ``count``; methods ``reset`` and ``update(val, n=1)``)."""


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
