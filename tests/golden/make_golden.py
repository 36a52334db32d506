"""How the fixtures in this directory were made (run by hand in the development container, where the
reference checkout is mounted at /root/reference; nothing at test or bench time reads that path).

  ../../maray_b200/data/chess.maray
                         copy of /root/reference/data/chess.maray   (sha256 b1ad82f4...baaba4, legacy wire layout);
                         it is an input scene of the benchmark, so the package owns it
  chess_reference.png    copy of /root/reference/images/chess.png    (decoded RGB8 sha256 b6f0efcf...2ccac2)
  chess_oracle_1024.png  the oracle's own render of chess.maray at its stored size
                         (decoded RGB8 sha256 4d2ca7dd...743bba -- the value an independent numpy evaluator
                         produced during the survey)

usage: python tests/golden/make_golden.py [/root/reference]
"""
import hashlib
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main() -> None:
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    scene_path = os.path.join(os.path.dirname(os.path.dirname(HERE)), "maray_b200", "data", "chess.maray")
    shutil.copyfile(os.path.join(ref, "data", "chess.maray"), scene_path)
    shutil.copyfile(os.path.join(ref, "images", "chess.png"), os.path.join(HERE, "chess_reference.png"))
    from oracle.oracle import OracleScene

    with open(scene_path, "rb") as f:
        raw = f.read()
    rgb = OracleScene(raw).render()
    Image.fromarray(rgb).save(os.path.join(HERE, "chess_oracle_1024.png"))
    print("chess.maray          ", hashlib.sha256(raw).hexdigest())
    print("chess_reference RGB8 ", hashlib.sha256(np.array(Image.open(os.path.join(HERE, "chess_reference.png")).convert("RGB")).tobytes()).hexdigest())
    print("chess_oracle RGB8    ", hashlib.sha256(rgb.tobytes()).hexdigest())


if __name__ == "__main__":
    main()
