#!/usr/bin/env python
"""Gradient-magnitude command line tool on the B200 path -- same arguments as the reference's ``bin/medpy_gradient.py``
(``input output [-v] [-d] [-f]``), written from scratch: Prewitt gradient magnitude as float32, the pre-step of
``medpy_graphcut_voxel.py --boundary max_*``."""
import argparse
import logging
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "compat"))
sys.path.insert(1, os.path.dirname(os.path.dirname(_HERE)))


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("input", help="Source volume.")
    p.add_argument("output", help="Target volume.")
    p.add_argument("-v", dest="verbose", action="store_true", help="Display more information.")
    p.add_argument("-d", dest="debug", action="store_true", help="Display debug information.")
    p.add_argument("-f", dest="force", action="store_true", help="Silently override existing output images.")
    args = p.parse_args(argv)
    from medpy.core import Logger
    from medpy.io import load, save
    from medpy_b200.gradient import gradient_magnitude_prewitt
    logger = Logger.getInstance()
    if args.debug:
        logger.setLevel(logging.DEBUG)
    elif args.verbose:
        logger.setLevel(logging.INFO)
    if not args.force and os.path.exists(args.output):
        logger.warning("The output image {} already exists. Exiting.".format(args.output))
        return -1
    data, hdr = load(args.input)
    logger.info("Computing the gradient magnitude with Prewitt operator...")
    out = gradient_magnitude_prewitt(data)
    save(out, args.output, hdr, args.force)
    logger.info("Successfully terminated.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
