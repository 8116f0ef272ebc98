#!/usr/bin/env python
"""Voxel graph-cut command line tool on the B200 path -- same arguments as the reference's
``bin/medpy_graphcut_voxel.py`` (positional ``sigma badditional markers output``; ``--boundary`` one of
diff_linear|diff_exp|diff_div|diff_pow|max_linear|max_exp|max_div|max_pow; ``-s`` use voxel spacing; ``-f`` overwrite;
``-v`` / ``-d`` verbosity), written from scratch.  Differences: the mask is read back in one bulk copy instead of a
``what_segment`` call per voxel, and the two linear terms get the 2-tuple they expect (the reference passes a 3-tuple
and fails)."""
import argparse
import logging
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "compat"))
sys.path.insert(1, os.path.dirname(os.path.dirname(_HERE)))

BOUNDARY = {
    "diff_linear": "boundary_difference_linear", "diff_exp": "boundary_difference_exponential",
    "diff_div": "boundary_difference_division", "diff_pow": "boundary_difference_power",
    "max_linear": "boundary_maximum_linear", "max_exp": "boundary_maximum_exponential",
    "max_div": "boundary_maximum_division", "max_pow": "boundary_maximum_power",
}


def get_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    p.add_argument("sigma", type=float, help="The sigma required for the boundary terms.")
    p.add_argument("badditional", help="The additional image required by the boundary term.")
    p.add_argument("markers", help="Image containing the foreground (=1) and background (=2) markers.")
    p.add_argument("output", help="The output image containing the segmentation.")
    p.add_argument("--boundary", default="diff_exp", choices=sorted(BOUNDARY), help="The boundary term to use.")
    p.add_argument("-s", dest="spacing", action="store_true", help="Take the voxel spacing of the image into account.")
    p.add_argument("-f", dest="force", action="store_true", help="Silently override existing files.")
    p.add_argument("-v", dest="verbose", action="store_true", help="Display more information.")
    p.add_argument("-d", dest="debug", action="store_true", help="Display debug information.")
    return p


def main(argv=None):
    args = get_parser().parse_args(argv)
    import numpy
    from medpy import graphcut
    from medpy.core import ArgumentError, Logger
    from medpy.graphcut.wrapper import split_marker
    from medpy.io import header, load, save
    logger = Logger.getInstance()
    if args.debug:
        logger.setLevel(logging.DEBUG)
    elif args.verbose:
        logger.setLevel(logging.INFO)
    if not args.force and os.path.exists(args.output):
        logger.warning("The output image {} already exists. Exiting.".format(args.output))
        return -1
    term = getattr(graphcut.energy_voxel, BOUNDARY[args.boundary])
    image, hdr = load(args.badditional)
    markers, _ = load(args.markers)
    fg, bg = split_marker(markers)
    if not (image.shape == fg.shape == bg.shape):
        raise ArgumentError("Not all of the supplied images are of the same shape.")
    spacing = header.get_voxel_spacing(hdr) if args.spacing else False
    term_args = (image, spacing) if args.boundary.endswith("linear") else (image, args.sigma, spacing)
    logger.info("Preparing the lattice graph on the GPU...")
    g = graphcut.graph_from_voxels(fg, bg, boundary_term=term, boundary_term_args=term_args)
    logger.info("Executing min-cut...")
    flow = g.maxflow()
    logger.debug("Maxflow is {}".format(flow))
    mask = g.get_mask().reshape(bg.shape).astype(numpy.bool_)
    save(mask, args.output, hdr, args.force)
    logger.info("Successfully terminated.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
