"""python -m opencafemol_b200 input.toml: the reference's main program (src/OpenAICG2Plus.cpp) on the B200 engine."""
import sys

from .simulator import main

sys.exit(main())
