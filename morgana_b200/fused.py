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

    def __call__(self, pred, target, n_frames, want_grad=True):
        """-> ``(loss, grad)``: the 0-dim total loss and d loss / d pred (``None`` unless `want_grad`).

        ``pred`` / ``target``: (B, T, total_dim) float32; the vuv column of ``pred`` is a probability and of ``target`` is
        0/1.  ``n_frames``: (B,) lengths.  Metric state is updated in place (see ``self.metrics``).
        """
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
