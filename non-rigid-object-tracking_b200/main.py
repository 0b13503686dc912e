#!/usr/bin/env python
"""Sequence driver with the command line of the reference's main.py (:77-88, :366-368):

    python main.py                      # config.yaml, prints the score
    python main.py CONFIG RESULT_FILE   # benchmark mode: writes "{mean IoU};{seconds}"

The per-frame masker work runs on the GPU (maskers/ -> libpcm_b200.so); see pcm/sequence.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from pcm.sequence import load_config, run_sequence  # noqa: E402


def main(argv):
    config_file, out = os.path.join(HERE, "config.yaml"), None
    if len(argv) > 2:
        config_file, out = argv[1], argv[2]
    elif len(argv) == 2:
        config_file = argv[1]
    cfg = load_config(config_file)
    res = run_sequence(cfg, device=int(os.environ.get("LOCAL_RANK", "0")), out_path=out, verbose=True)
    print("\nTotal benchmark score: %s" % res["mean_iou"])
    print("Total time consumed for tracking: %.2fs (%d frames, %d masker updates, tracker provider: %s; "
          "training %.2fs)" % (res["seconds"], res["n_frames"], res["n_updates"], res["tracker"], res["train_seconds"]))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
