"""Recipe for `oracle/_ref/`: the reference's own implementation of the decode path, staged so that it can travel.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it means staging the few modules the
path executes - UNMODIFIED, byte for byte - from where they lie under /root/reference into `oracle/_ref/infgen/`.
`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so it
reaches the GPU box like the built `.so`; there `bench.py --impl reference` runs the reference itself on the host cores
through `oracle/shims` (`cpu_baseline.kind = "reference"`), and falls back to the CPU port when the directory is absent.

    python -m oracle.make_ref          # run by __graft_entry__.build() when /root/reference is present

Files staged (what `InfGenAgentDecoder.inference` / `InfGenMapDecoder.forward` import):
"""
import filecmp
import os
import shutil
import sys

REFERENCE = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
FILES = (
    'infgen/modules/agent_decoder.py',      # InfGenAgentDecoder.inference (:1605-2389)
    'infgen/modules/layers.py',             # AttentionLayer, FourierEmbedding, MLPEmbedding, MLPLayer
    'infgen/modules/attr_tokenizer.py',     # Attr_Tokenizer.encode_pos / decode_pos / decode_heading
    'infgen/modules/map_decoder.py',        # InfGenMapDecoder.forward (row f1)
    'infgen/utils/func.py',                 # wrap_angle, angle_between_2d_vectors, weight_init
    'infgen/datasets/preprocess.py',        # AGENT_SHAPE, AGENT_TYPE (imported by agent_decoder.py:16)
)


def stage(verbose: bool = True) -> bool:
    """Copy the files when /root/reference is present; returns True when oracle/_ref is complete afterwards."""
    if os.path.isdir(os.path.join(REFERENCE, 'infgen', 'modules')):
        for rel in FILES:
            src, dst = os.path.join(REFERENCE, rel), os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
        if verbose:
            print(f'oracle/_ref: {len(FILES)} reference modules staged from {REFERENCE}')
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


if __name__ == '__main__':
    sys.exit(0 if stage() else 1)
