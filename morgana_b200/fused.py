"""Additive fused entry points (not in the reference): the whole per-batch objective of an acoustic model in ONE launch.

``LSTMAcousticModel.loss`` (reference models/RNN_SPSS.py:120-139) issues four metric accumulations and four masked
losses -- roughly 60 ATen kernels, 4 ``.item()`` syncs and a dozen (B, T, D) temporaries per batch.
:class:`AcousticObjective` computes the same numbers with one ``mg_masked_reduce`` launch of eight terms over the
(B, T, 187) prediction / target pair: the three ``mse`` terms and the ``bce`` term with their gradient w.r.t. the
prediction written in the same pass, and the four streaming metrics added straight into device-resident state.
Parity with the composition of the drop-in ops (and hence with the reference) is checked in the tests.
"""
import torch

from morgana_b200 import _lib, ops
from morgana_b200 import metrics as M


class AcousticObjective(object):
    r"""Loss (+ gradient) and metrics of ``models/RNN_SPSS.py`` on the column layout ``lf0 | vuv | mcep | bap``.

    Parameters
    ----------
    output_dims : dict
        Widths of the four streams, in the order produced by ``torch.split`` (RNN_SPSS.py:86-88); deltas included.
    mcep_static, bap_static : int
        Number of static coefficients of the mcep / bap streams ([static | delta | delta-delta] blocks).
    """
    def __init__(self, output_dims=None, mcep_static=60, bap_static=1):
        dims = output_dims or {'lf0': 3, 'vuv': 1, 'mcep': 180, 'bap': 3}
        self.widths = [dims[n] for n in ('lf0', 'vuv', 'mcep', 'bap')]
        self.starts = [sum(self.widths[:i]) for i in range(4)]
        self.total_dim = sum(self.widths)
        self.mcep_static, self.bap_static = mcep_static, bap_static
        self.metrics = {'LF0_RMSE_Hz': M.LF0Distortion(), 'VUV_accuracy': M.Mean(),
                        'MCEP_distortion': M.MelCepDistortion(), 'BAP_distortion': M.Distortion()}
        self.reset_state()

    def reset_state(self):
        for metric in self.metrics.values():
            metric.reset_state()
        self._records = None

    def _bind_metric_records(self, device):
        """Give the four metrics one shared block of device records and point their `sum` / `count` views at it."""
        self._records = ops.new_result_records(4, device)
        for record, metric in zip(self._records, self.metrics.values()):
            metric._record = record
            metric._integer = False
            metric.hidden = metric._hidden
            metric.sum = record.view(torch.float32)[ops.F32_SUM]
            metric.count = record.view(torch.float64)[ops.F64_COUNT]

    def _column_program(self, device):
        """One :class:`_lib.Column` per feature column: which loss slot (0-3) and metric slot (4-7) it feeds."""
        (s_lf0, s_vuv, s_mcep, s_bap), (w_lf0, w_vuv, w_mcep, w_bap) = self.starts, self.widths
        none = _lib.COL_NONE
        cols = [_lib.Column(none, 0, none, 0, none, 1, 0.) for _ in range(self.total_dim)]
        for slot, (kind, start, width) in enumerate([(_lib.RED_SQDIFF, s_lf0, w_lf0), (_lib.RED_SQDIFF, s_mcep, w_mcep),
                                                     (_lib.RED_SQDIFF, s_bap, w_bap), (_lib.RED_BCE, s_vuv, w_vuv)]):
            for c in range(start, start + width):
                cols[c].loss_kind, cols[c].loss_slot, cols[c].loss_weight = kind, slot, 0.25 / width
        # LF0_RMSE_Hz: static log-F0 column, exp fused, weighted by (vuv probability > 0.5)   (RNN_SPSS.py:122, 126)
        cols[s_lf0].metric_kind, cols[s_lf0].metric_slot, cols[s_lf0].mask_col = _lib.RED_SQDIFF_EXP, 4, s_vuv
        # VUV_accuracy: (target == (probability > 0.5))                                         (RNN_SPSS.py:127)
        cols[s_vuv].metric_kind, cols[s_vuv].metric_slot = _lib.RED_EQ, 5
        # MCEP_distortion: static coefficients 1 .. mcep_static-1                                (metrics.py:690-694)
        for c in range(s_mcep + 1, s_mcep + self.mcep_static):
            cols[c].metric_kind, cols[c].metric_slot = _lib.RED_SQDIFF, 6
        # BAP_distortion: per-frame Euclidean distance over the static band aperiodicities      (metrics.py:657-665)
        cols[s_bap].metric_kind, cols[s_bap].metric_slot, cols[s_bap].width = _lib.RED_ROOT_SQDIFF, 7, self.bap_static
        self._slot_dims = [w_lf0, w_mcep, w_bap, w_vuv, 1, 1, self.mcep_static - 1, 1]
        return ops.column_table(cols, device)

    def __call__(self, pred, target, n_frames, want_grad=True, grad_scale_dev=None, loss_records=None):
        """-> ``(loss, grad)``: the 0-dim total loss and d loss / d pred (``None`` unless `want_grad`).

        ``pred`` / ``target``: (B, T, total_dim) float32; the vuv column of ``pred`` is a probability and of ``target`` is
        0/1.  ``n_frames``: (B,) lengths.  Metric state is updated in place (see ``self.metrics``).
        One launch of the whole-row kernel (K4b); ``last_loss_records`` holds the four per-term losses (written into
        ``loss_records``, a (4, 48) uint8 device block, when the caller keeps a per-step log of them).
        """
        ops._require_cuda(pred, 'pred')
        ops._require_cuda(target, 'target')
        B, T, D = pred.shape
        if D != self.total_dim or tuple(target.shape) != (B, T, D):
            raise RuntimeError('expected (B, T, {}) prediction and target, got {} and {}'.format(
                self.total_dim, tuple(pred.shape), tuple(target.shape)))
        if pred.dtype != torch.float32 or target.dtype != torch.float32:
            raise TypeError('AcousticObjective takes float32 tensors')
        if self._records is None or self._records.device != pred.device:
            self._bind_metric_records(pred.device)
            self._cols = self._column_program(pred.device)
            self._slots = (_lib.Slot * 8)()
            for i, dim in enumerate(self._slot_dims):
                sl = self._slots[i]
                sl.D, sl.weight = dim, 0.25
                sl.in_total = int(i < 4)
                sl.accumulate = int(i >= 4)
                sl.per_frame = int(i in (4, 7))
                sl.weighted = int(i == 4)
                if i >= 4:
                    sl.result = self._records[i - 4].data_ptr()
        if loss_records is None:
            loss_records = ops.new_output_records(4, pred.device)
        for i in range(4):
            self._slots[i].result = loss_records[i].data_ptr()
        grad = torch.empty((B, T, D), dtype=torch.float32, device=pred.device) if want_grad else None
        with ops._device_of(pred):
            ops.masked_objective(pred, target, n_frames, self._cols, self._slots, grad=grad, grad_scale_dev=grad_scale_dev)
        self.last_loss_records = loss_records
        loss = loss_records[0].view(torch.float32)[ops.F32_TOTAL]
        return loss, grad

    def call_with_terms(self, pred, target, n_frames, want_grad=True):
        """The same objective through eight column-slice terms of the general kernel (K4/K5); kept for cross-checking."""
        ops._require_cuda(pred, 'pred')
        ops._require_cuda(target, 'target')
        B, T, D = pred.shape
        if D != self.total_dim or tuple(target.shape) != (B, T, D):
            raise RuntimeError('expected (B, T, {}) prediction and target, got {} and {}'.format(
                self.total_dim, tuple(pred.shape), tuple(target.shape)))
        if self._records is None or self._records.device != pred.device:
            self._bind_metric_records(pred.device)
        loss_records = ops.new_result_records(4, pred.device)
        grad = torch.empty_like(pred, memory_format=torch.contiguous_format) if want_grad else None

        def cols(t, start, width):
            return t[:, :, start:start + width] if t is not None else None

        (s_lf0, s_vuv, s_mcep, s_bap), (w_lf0, w_vuv, w_mcep, w_bap) = self.starts, self.widths
        with ops._device_of(pred):
            terms = []
            for i, (kind, start, width) in enumerate([(_lib.RED_SQDIFF, s_lf0, w_lf0), (_lib.RED_SQDIFF, s_mcep, w_mcep),
                                                      (_lib.RED_SQDIFF, s_bap, w_bap), (_lib.RED_BCE, s_vuv, w_vuv)]):
                terms.append(ops.make_term(kind, cols(pred, start, width), cols(target, start, width),
                                           result=loss_records[i], grad=cols(grad, start, width), grad_scale=0.25,
                                           flags=_lib.FLAG_IN_TOTAL))
            vuv_prob = cols(pred, s_vuv, 1)
            rec = self._records
            terms.append(ops.make_term(_lib.RED_SQDIFF_EXP, cols(target, s_lf0, 1), cols(pred, s_lf0, 1), m=vuv_prob,
                                       result=rec[0], accumulate=True, flags=_lib.FLAG_M_GT_HALF))
            terms.append(ops.make_term(_lib.RED_EQ, vuv_prob, cols(target, s_vuv, 1), result=rec[1], accumulate=True,
                                       flags=_lib.FLAG_A_GT_HALF))
            terms.append(ops.make_term(_lib.RED_SQDIFF, cols(target, s_mcep + 1, self.mcep_static - 1),
                                       cols(pred, s_mcep + 1, self.mcep_static - 1), result=rec[2], accumulate=True))
            terms.append(ops.make_term(_lib.RED_ROOT_SQDIFF, cols(target, s_bap, self.bap_static),
                                       cols(pred, s_bap, self.bap_static), result=rec[3], accumulate=True))
            ops.masked_reduce(terms, n_frames, B, T, pred.device)
        self.last_loss_records = loss_records
        loss = loss_records[0].view(torch.float32)[ops.F32_TOTAL]
        return loss, grad
