#!/usr/bin/env python
"""Stage 2 entry point with the FFDNet-colour denoiser (drop-in for the reference script of the same name)."""
from adaptivepnp_sci_b200.stage2_script import main

if __name__ == "__main__":
    main('ffdnet_color')
