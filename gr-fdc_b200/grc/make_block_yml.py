#!/usr/bin/env python
"""Writes the GNU Radio >= 3.8 GRC descriptors (FDC_*.block.yml) of the B200 drop-in blocks.

The block ids and the make templates are the reference's (grc/FDC_*.xml: <key> and <make>), so flowgraphs saved with the
3.7 XML descriptors keep loading; parameter ids are the make arguments.  Run: python gr-fdc_b200/grc/make_block_yml.py"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
YES_NO = ("enum", "True", ["True", "False"], ["Yes", "No"])


def P(pid, label, dtype="int", default=None, options=None, labels=None, attrs=None):
    return dict(id=pid, label=label, dtype=dtype, default=default, options=options, labels=labels, attrs=attrs)


def E(pid, label):
    return P(pid, label, *YES_NO)


IO_TYPES = P("type", "IO Type", "enum", "complex", ["complex", "float", "int", "short", "byte"], ["Complex", "Float", "Int", "Short", "Byte"],
             {"size": ["gr.sizeof_gr_complex", "gr.sizeof_float", "gr.sizeof_int", "gr.sizeof_short", "gr.sizeof_char"]})

def N01(pid, label, default="0", labels=("No", "Yes")):
    return P(pid, label, "enum", default, ["0", "1"], list(labels))


HIER = dict(
    id="FDC_FrequencyDomainChannelizer", label="Frequency Domain Channelizer",
    make="FDC.FrequencyDomainChannelizer(${type.size}, ${inpveclen}, ${blocksize}, ${relinvovl}, ${throughput_channels}, "
         "${activity_controlled_channels}, ${act_contr_threshold}, ${fs}, ${centerfrequency}, ${freqmode}, ${windowtype}, ${msgoutput}, ${fileoutput}, "
         "${outputpath}, ${threaded}, ${activity_detection_segments}, ${act_det_threshold}, ${minchandist}, ${act_det_deactivation_delay}, "
         "${minchanflankpuffer}, ${verbose}, ${pow_act_deactivation_delay}, ${pow_act_maxblocks}, ${act_det_maxblocks}, ${debug})",
    params=[P("type", "Input type", "enum", "0", ["0", "1"], ["Complex", "Float"], {"size": ["gr.sizeof_gr_complex", "gr.sizeof_float"], "tp": ["complex", "float"]}),
            P("blocksize", "FFT size", default=4096), P("inpveclen", "Input vector length", default=1), P("relinvovl", "Rel. inverse overlap", default=4),
            N01("threaded", "Threaded"), P("freqmode", "Frequency mode", "enum", "0", ["0", "1", "2"], ["Normalized", "Baseband + fs", "Centre frequency + fs"]),
            P("fs", "Sample rate", "float", 1.0), P("centerfrequency", "Centre frequency", "float", 0.0), N01("debug", "Debug spectrum output"),
            N01("msgoutput", "Message output", "1"), N01("fileoutput", "File output"), P("outputpath", "Output path", "string", ""),
            P("verbose", "Verbose", "enum", "0", ["0", "1", "2"], ["No log", "Console", "File"]),
            P("windowtype", "Window", "enum", "1", ["0", "1", "2"], ["Rectangular", "Hann", "Ramp"]),
            P("throughput_channels", "Throughput channels [[f, bw], ...]", "raw", "[]"),
            P("activity_controlled_channels", "Activity controlled channels [[f, bw], ...]", "raw", "[]"),
            P("act_contr_threshold", "Activity control threshold [dB]", "float", 6.0),
            P("pow_act_deactivation_delay", "Activity control deactivation delay", default=1), P("pow_act_maxblocks", "Activity control max. blocks", default=256),
            P("activity_detection_segments", "Detection segments [[start, stop], ...]", "raw", "[]"),
            P("act_det_threshold", "Detection threshold [dB]", "float", 6.0), P("minchandist", "Minimum channel distance", "float", 0.02),
            P("minchanflankpuffer", "Minimum channel puffer", "float", 0.2), P("act_det_deactivation_delay", "Detection deactivation delay", default=1),
            P("act_det_maxblocks", "Detection max. blocks", default=256)],
    inputs=[dict(domain="stream", dtype="${ type.tp }", vlen="${ inpveclen }")],
    outputs=[dict(domain="stream", id="debug", dtype="complex", vlen="${ blocksize }", optional="true", hide="${ debug == 0 }"),
             dict(domain="stream", dtype="complex", multiplicity="${ len(throughput_channels) }"),
             dict(domain="message", id="msgout", optional="true")],
    doc="Overlap-save, forward FFT and every fixed / activity-gated channel behind one block; on the B200 build the front end and all "
        "throughput channels are one fused GPU context (fdc_chan_*), the activity-gated blocks read the spectrum in device memory.")

BLOCKS = [
    dict(id="FDC_overlap_save", label="Overlap Save", make="FDC.overlap_save(${type.size}, ${outputlen}, ${overlaplen})",
         params=[IO_TYPES, P("outputlen", "Output length", default=4096), P("overlaplen", "Overlap length", default=1024)],
         asserts=["${ overlaplen >= 1 }", "${ outputlen > overlaplen }"],
         inputs=[dict(domain="stream", dtype="${ type }", vlen="${ outputlen - overlaplen }")],
         outputs=[dict(domain="stream", dtype="${ type }", vlen="${ outputlen }")],
         doc="Every output vector is the last overlaplen items of the previous vector followed by outputlen - overlaplen new items "
             "(zeros before the first).  CUDA implementation behind fdc_overlap_save_*."),
    dict(id="FDC_vector_cut_vxx", label="Vector Cut", make="FDC.vector_cut_vxx(${type.size}, ${veclen}, ${offset}, ${blocklen})",
         params=[IO_TYPES, P("veclen", "Vector length", default=4096), P("offset", "Offset", default=0), P("blocklen", "Block length", default=1024)],
         asserts=["${ offset >= 0 }", "${ offset + blocklen <= veclen }"],
         inputs=[dict(domain="stream", dtype="${ type }", vlen="${ veclen }")],
         outputs=[dict(domain="stream", dtype="${ type }", vlen="${ blocklen }")],
         doc="out = in[offset : offset + blocklen] of every input vector."),
    dict(id="FDC_phase_shifting_windowing_vcc", label="Phase Shifting Windowing",
         make="FDC.phase_shifting_windowing_vcc(${blocklen}, ${numphasestates}, ${shifts}, ${passbw}, ${stopbw}, ${windowtype})",
         params=[P("blocklen", "Block length", default=1024), P("numphasestates", "Phase states", default=4), P("shifts", "Shifts", default=1),
                 P("passbw", "Pass bandwidth", "float", 0.5), P("stopbw", "Stop bandwidth", "float", 0.75),
                 P("windowtype", "Window", "enum", "0", ["0", "1", "2"], ["Rectangular", "Hann", "Ramp"])],
         asserts=["${ stopbw >= passbw }"],
         inputs=[dict(domain="stream", dtype="complex", vlen="${ blocklen }")],
         outputs=[dict(domain="stream", dtype="complex", vlen="${ blocklen }")],
         doc="Frequency-domain filter mask times a per-block phase: block b is multiplied by table[(b * shifts) mod numphasestates]."),
    dict(id="FDC_PowerActivationChannel", label="Power Activation Channel",
         make="FDC.PowerActivationChannel(${blocklen}, ${cfreq}, ${bw}, ${relinvovl}, ${thresh}, ${maxblocks}, ${deactivation_delay}, ${msg}, "
              "${fileoutput}, ${path}, ${verbose}, ${ID})",
         params=[P("blocklen", "Block length", default=4096), P("cfreq", "Centre frequency", "float", 0.5), P("bw", "Bandwidth", "float", 0.1),
                 P("relinvovl", "Rel. inverse overlap", default=4), P("thresh", "Threshold [dB]", "float", 6.0), P("maxblocks", "Maximum blocks", default=256),
                 P("deactivation_delay", "Deactivation delay", default=1), E("msg", "Message output"), E("fileoutput", "File output"),
                 P("path", "Output path", "string", ""), P("verbose", "Verbose", default=0), P("ID", "ID", default=0)],
         inputs=[dict(domain="stream", dtype="complex", vlen="${ blocklen }")],
         outputs=[dict(domain="message", id="msgout", optional="true")],
         doc="Watches the power of one channel of the spectrum vectors and emits the extracted baseband bursts as PDUs."),
    dict(id="FDC_SegmentDetection", label="SegmentDetection",
         make="FDC.SegmentDetection(${ID}, ${blocklen}, ${relinvovl}, ${seg_start}, ${seg_stop}, ${thresh}, ${minchandist}, ${window_flank_puffer}, "
              "${maxblocks_to_emit}, ${channel_deactivation_delay}, ${messageoutput}, ${fileoutput}, ${path}, ${threads}, ${verbose})",
         params=[P("ID", "ID", default=0), P("blocklen", "Blocklen", default=4096), P("relinvovl", "Rel. inverse Overlap", default=4),
                 P("seg_start", "Segment Start", "float", 0.1), P("seg_stop", "Segment Stop", "float", 0.9), P("thresh", "Threshold [dB]", "float", 6.0),
                 P("minchandist", "Minimum channel distance", "float", 0.02), P("window_flank_puffer", "Minimum channel puffer", "float", 0.1),
                 P("maxblocks_to_emit", "Maximum Blocks", default=256), P("channel_deactivation_delay", "Channel deactivation delay", default=1),
                 E("messageoutput", "Message Output"), E("fileoutput", "File Output"), P("path", "Output path", "string", ""),
                 E("threads", "Threading"), P("verbose", "Verbose", default=0)],
         inputs=[dict(domain="stream", dtype="complex", vlen="${ blocklen }")],
         outputs=[dict(domain="message", id="msgout", optional="true")],
         doc="Detects carriers inside a frequency segment from block-to-block power edges and emits every detected burst as PDUs."),
    dict(id="FDC_activity_detection_channelizer_vcm", label="Activity Detection Channelizer",
         make="FDC.activity_detection_channelizer_vcm(${blocklen}, ${segments}, ${thresh}, ${relinvovl}, ${maxblocks}, ${message}, ${fileoutput}, "
              "${path}, ${threads}, ${minchandist}, ${channel_deactivation_delay}, ${window_flank_puffer}, ${verbose})",
         params=[P("blocklen", "Block length", default=4096), P("segments", "Segments", "raw", "[[0.1, 0.9]]"), P("thresh", "Threshold [dB]", "float", 6.0),
                 P("relinvovl", "Rel. inverse overlap", default=4), P("maxblocks", "Maximum blocks", default=256), E("message", "Message output"),
                 E("fileoutput", "File output"), P("path", "Output path", "string", ""), E("threads", "Threading"),
                 P("minchandist", "Minimum channel distance", "float", 0.02), P("channel_deactivation_delay", "Channel deactivation delay", default=1),
                 P("window_flank_puffer", "Minimum channel puffer", "float", 0.1), P("verbose", "Verbose", default=0)],
         inputs=[dict(domain="stream", dtype="complex", vlen="${ blocklen }")],
         outputs=[dict(domain="message", id="msgout", optional="true")],
         doc="Several detection segments in one block (predecessor of SegmentDetection)."),
]


def emit(b):
    out = ["id: %s" % b["id"], "label: %s" % b["label"], "category: '[FDC]'", "flags: [python, cpp]", "", "templates:", "  imports: import FDC",
           "  make: %s" % b["make"], "", "parameters:"]
    for p in b["params"]:
        out += ["-   id: %s" % p["id"], "    label: %s" % p["label"], "    dtype: %s" % p["dtype"]]
        if p["default"] is not None:
            out.append("    default: %s" % (("'%s'" % p["default"]) if p["dtype"] in ("string", "enum", "raw") else p["default"]))
        if p["options"]:
            out.append("    options: [%s]" % ", ".join(p["options"]))
            out.append("    option_labels: [%s]" % ", ".join(p["labels"]))
        if p["attrs"]:
            out.append("    option_attributes:")
            for k, v in p["attrs"].items():
                out.append("        %s: [%s]" % (k, ", ".join(v)))
    for key in ("inputs", "outputs"):
        out += ["", key + ":"]
        for port in b[key]:
            first = True
            for k, v in port.items():
                out.append(("-   " if first else "    ") + "%s: %s" % (k, v)); first = False
    if b.get("asserts"):
        out += ["", "asserts:"] + ["- %s" % a for a in b["asserts"]]
    out += ["", "documentation: |-", "    " + b["doc"], "", "file_format: 1", ""]
    with open(os.path.join(HERE, b["id"] + ".block.yml"), "w") as fh:
        fh.write("\n".join(out))


if __name__ == "__main__":
    for b in BLOCKS + [HIER]:
        emit(b)
    print("\n".join(sorted(f for f in os.listdir(HERE) if f.endswith(".yml"))))
