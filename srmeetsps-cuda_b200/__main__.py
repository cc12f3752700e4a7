"""CLI with the reference's flags (Main.cpp:10-17): --dstype/-t, --dsloc/-d, --device/-g,
--blockx/-x, --blocky/-y, --help/-h/--usage, `--key=value` syntax.  Extensions: --albedo=closed_form|reference_cg,
--outdir=DIR (result dumps and renderings, output.py)."""
import sys

from .srps import ImageDataHandler, MatFileDataHandler, Preferences, SRPS

HELP = """Usage: srmeetsps-cuda_b200 [params]

	-d, --dsloc
		path to dataset mat file or folder containing images
	-g, --device (value:0)
		cuda device to run the application on
	-h, --help, --usage
		print help
	-t, --dstype (value:matlab)
		dataset type, can be matlab or images
	-x, --blockx (value:256)
		block dimension x
	-y, --blocky (value:4)
		block dimension y
"""
ALIASES = {"t": "dstype", "d": "dsloc", "g": "device", "x": "blockx", "y": "blocky", "h": "help", "usage": "help"}


def parse(argv):
    opts = {"dstype": "matlab", "device": "0", "blockx": "256", "blocky": "4"}
    for a in argv:
        if not a.startswith("-"):
            continue
        key, _, val = a.lstrip("-").partition("=")
        key = ALIASES.get(key, key)
        opts[key] = val if val != "" else "true"
    return opts


def main(argv=None):
    opts = parse(sys.argv[1:] if argv is None else argv)
    if "help" in opts or "dsloc" not in opts:       # Main.cpp:19-26
        print(HELP)
        return 0
    Preferences.blockX = int(opts["blockx"])
    Preferences.blockY = int(opts["blocky"])
    Preferences.deviceId = int(opts["device"])
    if "albedo" in opts:
        Preferences.albedo_mode = opts["albedo"]
    if opts["dstype"] == "matlab":                   # Main.cpp:31-36
        dh = MatFileDataHandler().loadDataFromMatFiles(opts["dsloc"])
        SRPS(dh).execute(out_dir=opts.get("outdir"))
    elif opts["dstype"] == "images":                 # Main.cpp:37-42
        dh = ImageDataHandler().loadDataFromImages(opts["dsloc"])
        SRPS(dh).execute(out_dir=opts.get("outdir"))
    return 0                                         # any other dstype: silently nothing (Main.cpp:43)


if __name__ == "__main__":
    sys.exit(main())
