"""CPU numpy oracle of the data plumbing around the path.  TEST INFRASTRUCTURE ONLY — only tests/
may import it; the product (youreditableavatar_b200/formats.py) never does.

Restated (paths relative to /root/reference/Edit_core/):
  binding rule        tetgs_scene/tetgs_model.py:317-377 (area, mean_area, surface_n_gaussians,
                      calculate_attr_by_bary_coords) with utils/graphics_utils.py:118-123 (triangle_area)
  keep inheritance    tetgs_scene/tetgs_model.py:680-726 (convert_refined_tetgs_into_masked_gaussians: numpy
                      intersect1d / isin / where — restated literally)
  edit sub-mesh       tetgs_scene/tetgs_edit_2d.py:82-99
  2-D Gaussian frame  tetgs_scene/tetgs_edit_2d.py:172-208 with utils/graphics_utils.py:125-137 (calculate_distances)

Pinning: triangle_area and calculate_distances are pinned on outputs of the reference's own functions
(tests/golden/train_binding.npz, oracle/make_golden_train.py).  The methods of TetGS / EditTetGS themselves cannot be
imported here (pytorch3d, open3d, nerfstudio wrappers are absent and are used, not just imported), so the rule and the
inheritance are "parity unpinned" beyond those two functions and the literal restatement below.
"""
import numpy as np


def triangle_area(A, B, C):
    """graphics_utils.py:118-123"""
    return 0.5 * np.linalg.norm(np.cross(B - A, C - A), axis=1)


def bind_faces(verts, faces):
    """tetgs_model.py:328-377: one Gaussian at (1/3,1/3,1/3) on faces with area < mean area, else three at the
    permutations of (2/3,1/6,1/6); faces-with-one first, then faces-with-three (each repeated 3x)."""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    area = triangle_area(verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]])
    one = area < area.mean()
    f1, f3 = np.where(one)[0], np.where(~one)[0]
    b1 = np.tile(np.array([[1 / 3, 1 / 3, 1 / 3]]), (len(f1), 1))
    b3 = np.tile(np.array([[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]]), (len(f3), 1))
    return np.concatenate([f1, np.repeat(f3, 3)]), np.concatenate([b1, b3], 0)


def min_vertex_distance(points, A, B, C):
    """graphics_utils.py:125-137"""
    d = np.stack([np.linalg.norm(points - X, axis=1) for X in (A, B, C)], 0)
    return d.min(0)


def inherit_keep(face_indices, face_to_global_tet_idx, edit_face_to_global_tet_idx):
    """tetgs_model.py:696-702: indices of the Gaussians to keep."""
    face_mask = np.isin(face_to_global_tet_idx, edit_face_to_global_tet_idx)
    inherit_face_indices = np.where(face_mask)[0]
    gaussians_mask = np.isin(face_indices, inherit_face_indices)
    return np.where(gaussians_mask)[0]


def split_edit(vertices, faces, keep_vertices_num, keep_faces_num):
    """tetgs_edit_2d.py:94-98"""
    return vertices[keep_vertices_num:], faces[keep_faces_num:] - keep_vertices_num
