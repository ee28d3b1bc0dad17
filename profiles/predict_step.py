"""Fits the headline model once (crossPtDecoder_mcca, config 2) and calls FusedPredictor.predict
at batch 1 and batch 256 -- the launches profiled for BASELINE config 5."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sklearn.pipeline import make_pipeline
from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA
from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import crossPtDecoder_mcca
from cross_patient_speech_decoding_b200.decoders.fused_predict import FusedPredictor
from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
from cross_patient_speech_decoding_b200.folds import cv_splits
from cross_patient_speech_decoding_b200.svm import LinearSVC
pts = bench.make_data()
Xt, yt, yat = pts[0]
np.random.seed(0)
tr, te = cv_splits(yt, 20)[0]
m = crossPtDecoder_mcca(pts[1:], make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC()),
                        AlignMCCA, n_comp=30, regs=0.5, pca_var=0.8)
m.fit(Xt[tr], yt[tr], y_align=yat[tr])
fp = FusedPredictor(m)
X256 = np.ascontiguousarray(np.concatenate([Xt] * 2)[:256])
for _ in range(3):
    a = fp.predict(X256[:1])
    b = fp.predict(X256)
print('batch 1 ->', a, ' batch 256 agreement with predict():', float(np.mean(b == m.predict(X256))))
