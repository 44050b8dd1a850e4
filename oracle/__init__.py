"""CPU oracle for the learned-compression inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of hassanrizwank/Contextual-Image-Compression
for the path named in BASELINE.json (autoencoder, GAN encoder/generator codec, saliency-scaled
quantiser, ROI blend, bpp/PSNR/SSIM evaluation).  Only `tests/`, `__graft_entry__.smoke()` and
the CPU-baseline / `--impl reference` legs of `bench.py` may import it; the product package
never does and fails loudly when its CUDA library is missing.

PARITY UNPINNED (stated per the task contract): the reference ships no tests, golden vectors,
weights or datasets, and its arithmetic lives in TensorFlow/Keras 3 and scikit-image, neither
of which is installed or installable here (no network, not in /opt/wheelhouse; no version is
pinned by the reference either - it has no requirements file).  The oracle is therefore a
restatement of those libraries' published semantics at the reference's own
call sites, each function citing the reference file:line it follows.  What *is* pinned:
  * the closed-form known answers of SURVEY.md App. C (tests/golden/known_answers.json),
  * `cv2.cvtColor(BGR2GRAY)` - OpenCV 4.13 is importable, so the grayscale restatement is
    checked against the real library (tests/test_oracle_metrics.py),
  * `scipy.ndimage.uniform_filter`, which scikit-image's SSIM calls, is used directly,
  * an independent numpy direct-convolution restatement of TF 'same' padding checks the torch
    graphs (tests/test_oracle_graphs.py),
  * outputs of this oracle on seeded inputs are frozen under tests/golden/ by
    tests/golden/make_golden.py so that drift is caught.
"""
