"""Thin command-line shim for the three sub-commands on the hot path, with the reference's option names
(console.py:234-587, parse_config.py:391-660):

    python -m minorg_b200 grna       -t targets.fasta [-p .NGG] [-l 20] [-d out] [--prefix minorg]
    python -m minorg_b200 filter     --map out/minorg/minorg_gRNA_all.map -t targets.fasta -q genomeA.fasta -b bg.fasta
                                     [--assembly ref.fasta --screen-ref] [--ot-mismatch 1] [--ot-gap 0] [--ot-pattern P]
                                     [--ot-pam] [--gc-min .3 --gc-max .7] [--skip-bg-check] [-e exclude.fasta] [--mask m.fasta]
    python -m minorg_b200 minimumset --map out/minorg/minorg_gRNA_all.map [-s 1]
    python -m minorg_b200 full       = grna + filter + minimumset (targets given with -t; there is no seq() step)

Target discovery (`minorg seq`), GFF-based feature filtering and the interactive/manual minimum-set mode are not
part of this build."""
import argparse
import sys

from .grna import gRNAHits
from .MINORg import MINORg


def _common(p):
    p.add_argument("-d", "--dir", "--directory", dest="directory", default=None)
    p.add_argument("--prefix", default="minorg")
    p.add_argument("-t", "--target", default=None, help="FASTA of target sequences")
    p.add_argument("--device", type=int, default=None)


def _grna_opts(p):
    p.add_argument("-p", "--pam", default=".NGG")
    p.add_argument("-l", "--len", "--length", dest="length", type=int, default=20)


def _filter_opts(p):
    p.add_argument("-q", "--query", action="append", default=[])
    p.add_argument("-b", "--bg", "--background", dest="background", action="append", default=[])
    p.add_argument("--assembly", action="append", default=[], help="reference assembly FASTA (screened with --screen-ref)")
    p.add_argument("--screen-ref", "--screen-reference", dest="screen_reference", action="store_true")
    p.add_argument("--mask", action="append", default=[])
    p.add_argument("--ot-mismatch", type=int, default=1)
    p.add_argument("--ot-gap", type=int, default=0)
    p.add_argument("--ot-pattern", default=None)
    p.add_argument("--ot-unaligned-as-mismatch-unset", "--ot-uam-unset", dest="uam", action="store_false")
    p.add_argument("--ot-unaligned-as-gap", "--ot-uag", dest="uag", action="store_true")
    p.add_argument("--ot-pam", dest="ot_pamless", action="store_false", help="require a PAM next to an off-target hit")
    p.add_argument("--ot-pamless", dest="ot_pamless", action="store_true")
    p.set_defaults(ot_pamless=True, uam=True)
    p.add_argument("--gc-min", type=float, default=0.3)
    p.add_argument("--gc-max", type=float, default=0.7)
    p.add_argument("--skip-bg-check", action="store_true")
    p.add_argument("--skip-gc-check", action="store_true")
    p.add_argument("-e", "--exclude", default=None)


def _make(args):
    m = MINORg(directory=args.directory, prefix=args.prefix, device=args.device)
    m.target = args.target
    if hasattr(args, "length"):
        m.length = args.length
        m.pam = args.pam
    if hasattr(args, "ot_mismatch"):
        m.ot_mismatch, m.ot_gap, m.ot_pamless = args.ot_mismatch, args.ot_gap, args.ot_pamless
        m.ot_unaligned_as_mismatch, m.ot_unaligned_as_gap = args.uam, args.uag
        m.ot_pattern = args.ot_pattern
        m.gc_min, m.gc_max, m.exclude, m.mask = args.gc_min, args.gc_max, args.exclude, list(args.mask)
        m.screen_reference = args.screen_reference
        m.reference = {"ref_%d" % (i + 1): f for i, f in enumerate(args.assembly)}
        for f in args.query:
            m.add_query(f)
        for f in args.background:
            m.add_background(f)
    return m


def _load_map(m, path):
    m.grna_hits = gRNAHits()
    m.grna_hits.parse_from_mapping(path, targets=m.target)
    m.grna_map = path


def _do_filter(m, args):
    if not args.skip_bg_check:
        m.filter_background()
    if not args.skip_gc_check:
        m.filter_gc()
    m.filter_exclude()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="minorg_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("grna")
    _common(g); _grna_opts(g)
    f = sub.add_parser("filter", aliases=["check"])
    _common(f); _filter_opts(f)
    f.add_argument("-m", "--map", required=True)
    f.add_argument("--pam", "-p", default=".NGG")
    s = sub.add_parser("minimumset")
    _common(s)
    s.add_argument("-m", "--map", required=True)
    s.add_argument("-s", "--set", "--sets", dest="sets", type=int, default=1)
    s.add_argument("--prioritise-nr", dest="prioritise_nr", action="store_true")
    s.add_argument("--prioritise-pos", dest="prioritise_nr", action="store_false")
    s.add_argument("--manual", dest="manual", action="store_true", help="approve every gRNA set at the terminal")
    s.add_argument("--auto", dest="manual", action="store_false")
    a = sub.add_parser("full")
    _common(a); _grna_opts(a); _filter_opts(a)
    a.add_argument("-s", "--set", "--sets", dest="sets", type=int, default=1)
    a.add_argument("--prioritise-nr", dest="prioritise_nr", action="store_true")
    a.add_argument("--prioritise-pos", dest="prioritise_nr", action="store_false")
    a.add_argument("--manual", dest="manual", action="store_true", help="approve every gRNA set at the terminal")
    a.add_argument("--auto", dest="manual", action="store_false")
    args = ap.parse_args(argv)
    m = _make(args)
    m.prioritise_nr = bool(getattr(args, "prioritise_nr", False))
    if args.cmd == "grna":
        m.grna()
    elif args.cmd in ("filter", "check"):
        _load_map(m, args.map)
        m.pam = args.pam
        _do_filter(m, args)
    elif args.cmd == "minimumset":
        _load_map(m, args.map)
        m.minimumset(sets=args.sets, manual=bool(getattr(args, "manual", False)))
    else:
        m.grna()
        _do_filter(m, args)
        m.minimumset(sets=args.sets, manual=bool(getattr(args, "manual", False)))
    m.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
