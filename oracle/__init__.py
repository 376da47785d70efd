"""CPU oracle for the multi-modal-uncertainty hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in explicit torch-CPU tensor arithmetic (matmul / exp / sum; no
``nn.Module``, no fused ATen attention / layer-norm / cross-entropy ops), the algorithm of the
reference's data-parallel hot path so that the CUDA implementation can be checked against it
on machines where ``/root/reference`` does not exist.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import it.  The product package (``multi-modal-uncertainty_b200``) never does: it fails
loudly when the CUDA library is missing instead of falling back to anything here.

Parity status
-------------
* FLAVA fusion transformer, MIMO transformer, losses, ``acc``, AdamW, cosine schedule, batch
  shaping and the robustness sampling schedule are PINNED: ``tests/golden/*.pt`` were generated
  by importing the unmodified reference modules in the build container
  (``tests/golden/make_golden.py``) and ``tests/test_oracle_golden.py`` checks this oracle
  against them.
* Predictive entropy / expected entropy / mutual information / ECE / confidence histograms
  and guided / random modality dropout have NO reference implementation
  (SURVEY.md section 0): **parity unpinned** for those; the definitions in
  ``oracle/uncertainty.py`` are the repo's own, written out in fp64.
* MMBT: ``oracle/mmbt.py`` and ``oracle/image_encoder.py`` restate ``src/mmbt.py`` and are PINNED
  to goldens from the unmodified reference module (``tests/golden/make_golden_mmbt.py``).  The
  BERT layer arithmetic itself lives in the un-vendored, unpinned third-party
  ``pytorch_pretrained_bert``; ``oracle/bert_restated.py`` restates its published definitions so
  that the reference module can run here -- with respect to the absent package that part is
  **parity unpinned (third-party)**.  ViLT: not restated.
"""

from . import fusion, optim, shaping, uncertainty  # noqa: F401
