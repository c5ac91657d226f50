#!/usr/bin/env python
"""Stage 2 entry point with the FastDVDnet denoiser (drop-in for the reference script of the same name)."""
from adaptivepnp_sci_b200.stage2_script import main

if __name__ == "__main__":
    main('fastdvd_color')
