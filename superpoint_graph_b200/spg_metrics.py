"""Evaluation bookkeeping on the device: the step right after the path.

Reference: learning/metrics.py:8-79 (`ConfusionMatrix`) as driven by `eval` / `eval_final`
(learning/main.py:229-311): per batch the logits go to the host, `np.argmax`, a Python loop adds
every superpoint's label histogram to one column of the matrix.  Here the logits stay where they
are: one kernel (`spg_confusion_count`, csrc/loader.cu) takes the argmax, filters the unlabelled
nodes and accumulates an int64 [C, C] matrix in device memory; only the final matrix (C*C*8 bytes)
is read back.  Integer arithmetic: identical to the reference's counts.
"""
import numpy as np
import torch

from . import ops


class ConfusionMatrix(object):
    """Same interface as learning/metrics.py ConfusionMatrix (the getters return what the
    reference's return), plus `count_predicted_batch_device` for logits that are still on the GPU."""

    def __init__(self, number_of_labels=2, device=None):
        self.number_of_labels = number_of_labels
        self._host = np.zeros((number_of_labels, number_of_labels))
        self._dev = None
        self._counters = None
        self._device = device

    # ---- device accumulation -------------------------------------------------------------
    def count_predicted_batch_device(self, outputs, label_mode, label_vec, want_predictions=False):
        """outputs [n, C] float32 CUDA logits, label_mode [n] int64 (-100 = no ground truth),
        label_vec [n, C] int64 per-class point counts — `targets[:,0]`, `targets[:,2:]` of
        learning/main.py:246.  Equivalent to filter_valid + count_predicted_batch(tvec,
        argmax(o, 1)) (main.py:259-262).  Returns the predictions of all n nodes if asked."""
        if self._dev is None:
            C = self.number_of_labels
            self._dev = torch.zeros((C, C), dtype=torch.int64, device=outputs.device)
            self._counters = torch.zeros(2, dtype=torch.int64, device=outputs.device)
        return ops.confusion_count(outputs, label_mode, label_vec, self._dev, self._counters,
                                   want_predictions)

    def accuracy(self):
        """Top-1 accuracy in percent over the labelled nodes seen by the device path — what
        `tnt.meter.ClassErrorMeter(accuracy=True).value()[0]` reports in main.py:239,261,264."""
        if self._counters is None:
            return 0
        n, ok = [int(v) for v in self._counters.cpu()]
        return 100.0 * ok / n if n > 0 else 0

    @property
    def confusion_matrix(self):
        if self._dev is None:
            return self._host
        return self._host + self._dev.cpu().numpy().astype(np.float64)

    # ---- host interface of the reference ---------------------------------------------------
    def count_predicted(self, ground_truth, predicted, number_of_added_elements=1):
        self._host[ground_truth][predicted] += number_of_added_elements

    def count_predicted_batch(self, ground_truth_vec, predicted):
        np.add.at(self._host.T, np.asarray(predicted), np.asarray(ground_truth_vec))

    def count_predicted_batch_hard(self, ground_truth_vec, predicted):
        np.add.at(self._host, (np.asarray(ground_truth_vec), np.asarray(predicted)), 1)

    def get_count(self, ground_truth, predicted):
        return self.confusion_matrix[ground_truth][predicted]

    def get_confusion_matrix(self):
        return self.confusion_matrix

    def get_intersection_union_per_class(self):
        cm = self.confusion_matrix
        diag = np.diag(cm)
        divisor = cm.sum(1) + cm.sum(0) - diag
        divisor = np.where(diag == 0, 1, divisor)
        return [float(d) / v for d, v in zip(diag, divisor)]

    def get_overall_accuracy(self):
        cm = self.confusion_matrix
        total = cm.sum()
        return float(np.trace(cm)) / (total if total != 0 else 1)

    def get_average_intersection_union(self):
        cm = self.confusion_matrix
        values = self.get_intersection_union_per_class()
        class_seen = ((cm.sum(1) + cm.sum(0)) != 0).sum()
        return sum(values) / class_seen

    def get_mean_class_accuracy(self):
        cm = self.confusion_matrix
        re = 0
        for i in range(self.number_of_labels):
            re = re + cm[i][i] / max(1, np.sum(cm[i, :]))
        return re / self.number_of_labels

    def count_gt(self, ground_truth):
        return self.confusion_matrix[ground_truth, :].sum()


class MultiSampleMean(object):
    """`np.mean(np.stack(o_cpu, 0), 0)` of eval_final (learning/main.py:292-295) without leaving
    the device: numpy reduces axis 0 by adding the samples in order in float32 and divides once;
    the same two elementwise operations, so the averaged logits (and their argmax) are identical."""

    def __init__(self):
        self._sum, self._n = None, 0

    def add(self, outputs):
        self._sum = outputs.detach().clone() if self._sum is None else self._sum.add_(outputs)
        self._n += 1

    def value(self):
        if self._n == 1:
            return self._sum
        # tensor / tensor: IEEE division as numpy's true_divide (torch's CUDA tensor / python-scalar
        # multiplies by the reciprocal, which differs in the last bit)
        return self._sum / torch.full_like(self._sum, float(self._n))


# ---- label up-sampling (the step after the metrics: predictions of the pruned cloud -> full cloud) ----
def reduced_labels2full(labels_red, components, n_ver, device=None):
    """labels of superpoints -> labels of their member points (ref: partition/provider.py:630-635).
    `components`: list of index arrays (as read from the SPG file).  Returns a uint8 CUDA tensor."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    sizes = np.asarray([len(c) for c in components], dtype=np.int64)
    ptr = np.zeros(len(components) + 1, dtype=np.int64)
    np.cumsum(sizes, out=ptr[1:])
    ids = (np.concatenate([np.asarray(c, dtype=np.int64).reshape(-1) for c in components])
           if len(components) else np.zeros(0, dtype=np.int64))
    lab = torch.as_tensor(np.asarray(labels_red), dtype=torch.int64).reshape(-1).to(device)
    return ops.labels_to_points(lab, torch.from_numpy(ptr).to(device), torch.from_numpy(ids).to(device), int(n_ver))


def interpolate_labels(xyz_up, xyz, labels, ver_batch=0, return_index=False):
    """Labels of the pruned cloud `xyz` [n,3] -> full cloud `xyz_up` [m,3] by exact 1-nearest
    neighbour (ref: partition/provider.py:676-682; the reference's kd-tree works in float64 on the
    float32 coordinates, and so does the kernel).  `labels` [n] or [n,C] (argmax taken, as the
    reference does).  `ver_batch` > 0 processes the queries in slices of that many points
    (provider.py:637-675 reads the full cloud in batches).  Returns int64 CUDA labels [m]."""
    dev = xyz_up.device if torch.is_tensor(xyz_up) and xyz_up.is_cuda else torch.device("cuda", torch.cuda.current_device())
    ref = torch.as_tensor(xyz, dtype=torch.float32).to(dev).contiguous()
    qry = torch.as_tensor(xyz_up, dtype=torch.float32).to(dev).contiguous()
    lab = torch.as_tensor(labels).to(dev)
    if lab.dim() > 1 and lab.shape[1] > 1:
        lab = torch.argmax(lab, dim=1)
    lab = lab.reshape(-1).to(torch.int64).contiguous()
    step = int(ver_batch) if ver_batch and ver_batch > 0 else qry.shape[0]
    outs, idxs = [], []
    for i in range(0, qry.shape[0], max(step, 1)):
        o, ix = ops.nn1_interpolate(ref, qry[i:i + step], lab, want_index=return_index)
        outs.append(o)
        idxs.append(ix)
    out = torch.cat(outs) if outs else torch.zeros(0, dtype=torch.int64, device=dev)
    if return_index:
        return out, (torch.cat(idxs) if idxs else torch.zeros(0, dtype=torch.int32, device=dev))
    return out
