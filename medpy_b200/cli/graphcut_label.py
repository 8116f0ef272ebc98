#!/usr/bin/env python
"""Region (label) graph-cut command line tool on the B200 path -- same arguments as the reference's
``bin/medpy_graphcut_label.py`` (positional ``badditional region markers output``; ``--boundary`` stawiaski|means;
``-f`` overwrite; ``-v`` / ``-d`` verbosity), written from scratch.  ``badditional`` is the gradient magnitude image
for ``stawiaski`` and the original image for ``means``.  Differences: the region adjacency graph is reduced on the GPU
instead of one Python call per border voxel pair, and the cut is mapped back onto the voxels by one device gather
instead of ``what_segment`` per region + ``relabel_map``."""
import argparse
import logging
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "compat"))
sys.path.insert(1, os.path.dirname(os.path.dirname(_HERE)))


def get_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    p.add_argument("badditional", help="The additional image required by the boundary term.")
    p.add_argument("region", help="The region (label) image of the image to segment.")
    p.add_argument("markers", help="Image containing the foreground (=1) and background (=2) markers.")
    p.add_argument("output", help="The output image containing the segmentation.")
    p.add_argument("--boundary", default="stawiaski", choices=["means", "stawiaski"], help="The boundary term to use.")
    p.add_argument("-f", dest="force", action="store_true", help="Silently override existing files.")
    p.add_argument("-v", dest="verbose", action="store_true", help="Display more information.")
    p.add_argument("-d", dest="debug", action="store_true", help="Display debug information.")
    return p


def main(argv=None):
    args = get_parser().parse_args(argv)
    import numpy
    from medpy import filter, graphcut
    from medpy.core import ArgumentError, Logger
    from medpy.graphcut.wrapper import split_marker
    from medpy.io import load, save
    logger = Logger.getInstance()
    if args.debug:
        logger.setLevel(logging.DEBUG)
    elif args.verbose:
        logger.setLevel(logging.INFO)
    if not args.force and os.path.exists(args.output):
        logger.warning("The output image {} already exists. Exiting.".format(args.output))
        return -1
    term = graphcut.energy_label.boundary_stawiaski if args.boundary == "stawiaski" else graphcut.energy_label.boundary_difference_of_means
    regions, hdr = load(args.region)
    additional, _ = load(args.badditional)
    markers, _ = load(args.markers)
    fg, bg = split_marker(markers)
    if not (additional.shape == regions.shape == fg.shape == bg.shape):
        raise ArgumentError("Not all of the supplied images are of the same shape.")
    logger.info("Relabel input image...")
    regions = filter.relabel(regions)
    logger.info("Preparing the region graph on the GPU...")
    g = graphcut.graph_from_labels(regions, fg, bg, boundary_term=term, boundary_term_args=additional)
    logger.info("Executing min-cut...")
    flow = g.maxflow()
    logger.debug("Maxflow is {}".format(flow))
    mask = graphcut.label_cut_mask(g).astype(numpy.bool_)
    save(mask, args.output, hdr, args.force)
    logger.info("Successfully terminated.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
