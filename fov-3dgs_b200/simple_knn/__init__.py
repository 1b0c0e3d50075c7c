"""Drop-in for the reference's `simple_knn` package (fov3dgs/submodules/simple-knn): `from simple_knn._C import distCUDA2`
(scene/gaussian_model.py:20) resolves to the hash-grid 3-NN kernel of libfovgs.so."""
