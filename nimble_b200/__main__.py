"""`python -m nimble_b200 <subcommand>` — same sub-commands and flags as `python -m nimble`
(nimble/__main__.py:373-468) for the hot path: generate, align, report.  `download` reports that
the aligner is built in; `fastq-to-bam` (nimble/__main__.py:424-431) runs its barcode correction on
the GPU; `plot` is outside the hot path."""
import argparse
import sys

from . import __version__
from .frontend import align, align_10x, fastq_to_bam_with_barcodes, generate, report


def main(argv=None):
    parser = argparse.ArgumentParser(description="nimble align (B200-native backend)")
    parser.add_argument("-v", "--version", action="version", version=f"nimble_b200 {__version__}")
    sub = parser.add_subparsers(title="subcommands", dest="subcommand")

    d = sub.add_parser("download")
    d.add_argument("--release", help="The release to download.", type=str, default=[])

    g = sub.add_parser("generate")
    g.add_argument("--file", help="The file to process.", type=str, required=True)
    g.add_argument("--opt-file", help="The optional file to process.", type=str, default=None)
    g.add_argument("--output_path", help="The path to the output file.", type=str, required=True)

    a = sub.add_parser("align")
    a.add_argument("--reference", help="The reference genome to align to.", type=str, required=True)
    a.add_argument("--output", help="The path to the output file.", type=str, required=True)
    a.add_argument("--input", help="The input reads.", type=str, required=True, nargs="+")
    a.add_argument("-c", "--num_cores", help="The number of cores to use for alignment.", type=int, default=1)
    a.add_argument("--strand_filter", help="Filter reads based on strand information.", type=str, default="unstranded")
    a.add_argument("--trim", help="<TARGET_LENGTH>:<STRICTNESS>, comma-separated, one entry per library", type=str, default="")
    a.add_argument("--tmpdir", help="Path to a temporary directory for sorting .bam files", type=str, default=None)
    a.add_argument("-k", "--kmer", help="k-mer length of the index (4..32)", type=int, default=20)
    a.add_argument("--map", help="(extension) cell barcode whitelist: --input is a raw 10x R1/R2 FASTQ pair, fastq-to-bam runs in "
                                 "the same pass without writing a BAM", type=str, default=None)
    a.add_argument("--gpus", help="(extension) GPUs of this node to spread the reads over (default 1; also $NB200_GPUS)", type=int, default=None)
    a.add_argument("--cb-length", type=int, default=16)
    a.add_argument("--umi-length", type=int, default=12)

    r = sub.add_parser("report")
    r.add_argument("-i", "--input", help="The input file.", type=str, required=True)
    r.add_argument("-o", "--output", help="The path to the output file.", type=str, required=True)
    r.add_argument("-s", "--summarize", help="CSV list of columns to summarize.", type=str, default=None)
    r.add_argument("-t", "--threshold", help="Proportional count threshold for filtering features (default: 0.05).",
                   type=float, default=0.05)
    r.add_argument("--disable_thresholding", help="Disable the per-UMI proportional count thresholding algorithm.",
                   action="store_true", default=False)

    f = sub.add_parser("fastq-to-bam")
    f.add_argument("--r1-fastq", help="Path to R1 FASTQ file.", type=str, required=True)
    f.add_argument("--r2-fastq", help="Path to R2 FASTQ file.", type=str, required=True)
    f.add_argument("--map", required=True, help="Cell barcode whitelist file (one CB per line, .gz or plain text)")
    f.add_argument("--output", help="Path for output BAM file.", type=str, required=True)
    f.add_argument("-c", "--num_cores", help="The number of cores to use for processing.", type=int, default=1)
    f.add_argument("--cb-length", help="Length of cell barcode (default: 16).", type=int, default=16)
    f.add_argument("--umi-length", help="Length of UMI (default: 12).", type=int, default=12)

    pl = sub.add_parser("plot")          # visualisation: outside the hot path; flags accepted so scripts fail clearly
    pl.add_argument("--input_file", "-i", type=str, required=False)
    pl.add_argument("--output_file", "-o", type=str, required=False)

    args = parser.parse_args(argv)
    if args.subcommand == "download":
        print("nimble_b200: the aligner is the in-tree CUDA library (libnimble_b200.so); nothing to download.")
    elif args.subcommand == "generate":
        generate(args.file, args.opt_file, args.output_path)
    elif args.subcommand == "align" and args.map:
        if len(args.input) != 2:
            parser.error("--map needs exactly two --input files (R1 and R2 FASTQ)")
        if args.trim:
            print("nimble_b200: --trim is applied by the streaming file path only; ignored with --map "
                  "(run fastq-to-bam, then align --trim on its BAM)", file=sys.stderr)
        sys.exit(align_10x(args.reference, args.output, args.input[0], args.input[1], args.map, args.num_cores, args.strand_filter,
                           args.cb_length, args.umi_length, k=args.kmer))
    elif args.subcommand == "align":
        sys.exit(align(args.reference, args.output, args.input, args.num_cores, args.strand_filter, args.trim,
                       args.tmpdir, k=args.kmer, gpus=args.gpus))
    elif args.subcommand == "report":
        cols = args.summarize.split(",") if args.summarize else None
        report(args.input, args.output, cols, args.threshold, args.disable_thresholding)
    elif args.subcommand == "plot":
        print("nimble_b200: `plot` (HTML report, nimble/report_generation.py) is not part of the B200 backend; the per-read TSV "
              "written by `align` carries every column it reads, so `python -m nimble plot` works on it unchanged.", file=sys.stderr)
        sys.exit(2)
    elif args.subcommand == "fastq-to-bam":
        fastq_to_bam_with_barcodes(args.r1_fastq, args.r2_fastq, args.map, args.output, args.num_cores,
                                   args.cb_length, args.umi_length)
    else:
        parser.print_help()


if __name__ == "__main__":
    main()
