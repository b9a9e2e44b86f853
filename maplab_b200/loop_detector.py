"""Python mirror of matching_based_loopclosure::LoopDetector
(matching-based-loopclosure/include/matching-based-loopclosure/matching-based-engine.h:18-52)
for parity tests that read like the reference's own: same method names, argument meaning and
error behaviour (a violated precondition raises where the reference CHECK-aborts)."""
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class ProjectedImage:
    """loop_closure::ProjectedImage (descriptor-projection/descriptor-projection.h:23-31); ids dense."""
    timestamp_nanoseconds: int
    vertex_id: int
    frame_index: int
    mission_id: int
    projected_descriptors: np.ndarray            # [n][dim] float32
    landmarks: np.ndarray = None                 # [n] int64 (database images)
    measurements: np.ndarray = field(default=None)  # [n][2] keypoints (carried, unused here)


class LoopDetector:
    def __init__(self, vocabulary_blob, settings=None):
        self._d = capi.Detector(vocabulary_blob, settings)

    @property
    def detector(self):
        return self._d

    def Initialize(self):
        self._d.initialize()

    def Clear(self):
        self._d.clear()

    def NumEntries(self):
        return self._d.num_entries()

    def NumDescriptors(self):
        return self._d.num_descriptors()

    def ProjectDescriptors(self, descriptors):
        """descriptors: [n][bytes] uint8 (one per row) -> [n][dim] float32."""
        return self._d.project(descriptors)

    def Insert(self, image: ProjectedImage):
        n = len(image.projected_descriptors)
        if image.landmarks is not None and len(image.landmarks) != n:
            raise capi.MlcError("Insert: projected_descriptors.cols() != landmarks.size()")
        self._d.insert(image.timestamp_nanoseconds, image.vertex_id, image.frame_index,
                       image.mission_id, image.projected_descriptors, image.landmarks)

    def AddLocalizationSummaryMapToDatabase(self, file_bytes, mission_id, first_vertex_id, first_landmark_id):
        """LoopDetectorNode::addLocalizationSummaryMapToDatabase (LCH/src/loop-detector-node.cc:341-432)
        on the bytes of a `localization_summary_map` file."""
        return self._d.add_summary_map(file_bytes, mission_id, first_vertex_id, first_landmark_id)

    def Find(self, images, parallelize_if_possible=False):
        """images: ProjectedImage list of ONE vertex. Returns matches (capi.MATCH_DTYPE) in
        canonical order (query frame index, keypoint, db descriptor)."""
        if not images:
            return np.zeros(0, capi.MATCH_DTYPE)
        if any(im.vertex_id != images[0].vertex_id for im in images):
            raise capi.MlcError("Find: all images must belong to the same vertex")
        frames = capi.make_frames([im.timestamp_nanoseconds for im in images],
                                  [im.vertex_id for im in images],
                                  [im.mission_id for im in images],
                                  [im.frame_index for im in images],
                                  [len(im.projected_descriptors) for im in images])
        proj = np.concatenate([np.asarray(im.projected_descriptors, np.float32).reshape(-1, self._d.dim)
                               for im in images])
        matches, _ = self._d.find_batch(frames, proj=proj)
        return matches
