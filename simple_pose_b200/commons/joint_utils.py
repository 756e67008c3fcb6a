"""Flip semantics shared by the encoder-side augmentation and the flip-test decoder.

Mirror of the two pieces of the reference's ``commons/joint_utils.py`` that the heatmap hot
path consumes: the left/right pair swap of ``flip_joints`` (:102-112) expressed as a channel
permutation, with COCO's ``joint_pairs`` (``datasets/coco.py:26``) as the default.
"""
COCO_JOINT_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def swap_permutation(num_joints, joint_pairs=None):
    """perm such that ``flipped[k] = original[perm[k]]`` after the pair swap."""
    pairs = COCO_JOINT_PAIRS if joint_pairs is None else joint_pairs
    perm = list(range(num_joints))
    for a, b in pairs:
        if not (0 <= a < num_joints and 0 <= b < num_joints):
            raise ValueError("joint pair (%d, %d) outside 0..%d" % (a, b, num_joints - 1))
        perm[a], perm[b] = perm[b], perm[a]
    return perm
