#!/usr/bin/env python
"""Hand-assembled indexedmzML in the layout ProteoWizard's msconvert writes (PSI mzML 1.1.0 schema examples: cvList,
referenceableParamGroupList, pretty-printed elements, unit attributes on cvParams, scanWindowList, precursorList, a
chromatogramList behind the spectra, an indexList with a spectrum AND a chromatogram index, fileChecksum).

    python tests/data/make_pwiz_style.py   ->  tests/data/pwiz_style_small.mzML + pwiz_style_small_expected.npz

The binary arrays are encoded here with struct / zlib / base64 only -- NOT with pyqms_b200.mzml.write_mzml -- so the
readers are checked against a document their own writer did not produce.  No third-party converter is available in the
build container (no network); this is the "hand-assembled from the PSI spec examples" fixture.
"""
import base64
import struct
import zlib
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def encode(values, fmt, compress):
    raw = struct.pack("<%d%s" % (len(values), fmt), *values)
    if compress:
        raw = zlib.compress(raw, 6)
    return base64.b64encode(raw).decode()


def array(values, bits, compress, accession, name, unit, indent):
    text = encode(values, "d" if bits == 64 else "f", compress)
    pad = " " * indent
    rows = [pad + '<binaryDataArray encodedLength="%d">' % len(text),
            pad + '  <cvParam cvRef="MS" accession="%s" name="%d-bit float" value=""/>' % ("MS:1000523" if bits == 64 else "MS:1000521", bits),
            pad + '  <cvParam cvRef="MS" accession="%s" name="%s" value=""/>' % (("MS:1000574", "zlib compression") if compress else
                                                                                 ("MS:1000576", "no compression")),
            pad + '  <cvParam cvRef="MS" accession="%s" name="%s" value="" %s/>' % (accession, name, unit),
            pad + "  <binary>%s</binary>" % text if text else pad + "  <binary/>",
            pad + "</binaryDataArray>"]
    return "\n".join(rows)


MZ_UNIT = 'unitCvRef="MS" unitAccession="MS:1000040" unitName="m/z"'
I_UNIT = 'unitCvRef="MS" unitAccession="MS:1000131" unitName="number of detector counts"'
T_UNIT = 'unitCvRef="UO" unitAccession="UO:0000031" unitName="minute"'


def document(n_spectra=9):
    """(document bytes, expected arrays) of a run of n_spectra spectra."""
    rng = np.random.default_rng(20261020)
    spectra, levels, rts = [], [], []
    for n in range(n_spectra):
        k = [7, 0, 33, 5, 120, 1, 16, 64, 3][n % 9]
        mz = np.sort(rng.uniform(200.0, 1800.0, k))
        inten = np.float32(10 ** rng.uniform(2, 6, k))          # stored as 32-bit floats: exact after the round trip
        spectra.append(np.stack([mz, inten.astype(np.float64)], axis=1) if k else np.zeros((0, 2)))
        levels.append(2 if n % 9 in (3, 6) else 1)
        rts.append(round(0.5 + 0.0123 * n, 6))
    out, offsets, chroma_offset = [], [], None
    size = 0

    def emit(text):
        nonlocal size
        out.append(text)
        size += len(text.encode("utf-8"))

    emit('<?xml version="1.0" encoding="utf-8"?>\n<indexedmzML xmlns="http://psi.hupo.org/ms/mzml" '
         'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" xsi:schemaLocation="http://psi.hupo.org/ms/mzml '
         'http://psidev.info/files/ms/mzML/xsd/mzML1.1.2_idx.xsd">\n')
    emit('  <mzML xmlns="http://psi.hupo.org/ms/mzml" xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" '
         'xsi:schemaLocation="http://psi.hupo.org/ms/mzml http://psidev.info/files/ms/mzML/xsd/mzML1.1.0.xsd" id="small" version="1.1.0">\n')
    emit('    <cvList count="2">\n      <cv id="MS" fullName="Proteomics Standards Initiative Mass Spectrometry Ontology" version="4.1.30" '
         'URI="https://raw.githubusercontent.com/HUPO-PSI/psi-ms-CV/master/psi-ms.obo"/>\n      <cv id="UO" fullName="Unit Ontology" '
         'version="09:04:2014" URI="https://raw.githubusercontent.com/bio-ontology-research-group/unit-ontology/master/unit.obo"/>\n    </cvList>\n')
    emit('    <fileDescription>\n      <fileContent>\n        <cvParam cvRef="MS" accession="MS:1000579" name="MS1 spectrum" value=""/>\n'
         '        <cvParam cvRef="MS" accession="MS:1000580" name="MSn spectrum" value=""/>\n      </fileContent>\n    </fileDescription>\n')
    emit('    <referenceableParamGroupList count="1">\n      <referenceableParamGroup id="CommonInstrumentParams">\n'
         '        <cvParam cvRef="MS" accession="MS:1000448" name="LTQ FT" value=""/>\n      </referenceableParamGroup>\n'
         '    </referenceableParamGroupList>\n')
    emit('    <softwareList count="1">\n      <software id="pwiz" version="3.0.0">\n        <cvParam cvRef="MS" accession="MS:1000615" '
         'name="ProteoWizard software" value=""/>\n      </software>\n    </softwareList>\n')
    emit('    <instrumentConfigurationList count="1">\n      <instrumentConfiguration id="IC1">\n        '
         '<referenceableParamGroupRef ref="CommonInstrumentParams"/>\n      </instrumentConfiguration>\n    </instrumentConfigurationList>\n')
    emit('    <dataProcessingList count="1">\n      <dataProcessing id="pwiz_Reader_conversion">\n        <processingMethod order="0" '
         'softwareRef="pwiz">\n          <cvParam cvRef="MS" accession="MS:1000544" name="Conversion to mzML" value=""/>\n        '
         '</processingMethod>\n      </dataProcessing>\n    </dataProcessingList>\n')
    emit('    <run id="small" defaultInstrumentConfigurationRef="IC1" startTimeStamp="2026-10-20T09:00:00Z">\n')
    emit('      <spectrumList count="%d" defaultDataProcessingRef="pwiz_Reader_conversion">\n' % len(spectra))
    for n, (spectrum, level, rt) in enumerate(zip(spectra, levels, rts)):
        emit("        ")
        offsets.append(size)
        rows = ['<spectrum index="%d" id="controllerType=0 controllerNumber=1 scan=%d" defaultArrayLength="%d">' % (n, 19 + n, len(spectrum)),
                '          <cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="%d"/>' % level,
                '          <cvParam cvRef="MS" accession="%s" value=""/>' % ('MS:1000579" name="MS1 spectrum' if level == 1 else 'MS:1000580" name="MSn spectrum'),
                '          <cvParam cvRef="MS" accession="MS:1000127" name="centroid spectrum" value=""/>',
                '          <cvParam cvRef="MS" accession="MS:1000285" name="total ion current" value="%.6g"/>' % float(spectrum[:, 1].sum()),
                '          <scanList count="1">', '            <cvParam cvRef="MS" accession="MS:1000795" name="no combination" value=""/>',
                '            <scan instrumentConfigurationRef="IC1">',
                '              <cvParam cvRef="MS" accession="MS:1000016" name="scan start time" value="%r" %s/>' % (rt, T_UNIT),
                '              <cvParam cvRef="MS" accession="MS:1000512" name="filter string" value="FTMS + p ESI Full ms [200.00-2000.00]"/>',
                '              <scanWindowList count="1">', '                <scanWindow>',
                '                  <cvParam cvRef="MS" accession="MS:1000501" name="scan window lower limit" value="200.0" %s/>' % MZ_UNIT,
                '                  <cvParam cvRef="MS" accession="MS:1000500" name="scan window upper limit" value="2000.0" %s/>' % MZ_UNIT,
                '                </scanWindow>', '              </scanWindowList>', '            </scan>', '          </scanList>']
        if level == 2:
            rows += ['          <precursorList count="1">', '            <precursor spectrumRef="controllerType=0 controllerNumber=1 scan=%d">' % (19 + n - 1),
                     '              <selectedIonList count="1">', '                <selectedIon>',
                     '                  <cvParam cvRef="MS" accession="MS:1000744" name="selected ion m/z" value="810.79" %s/>' % MZ_UNIT,
                     '                </selectedIon>', '              </selectedIonList>', '            </precursor>', '          </precursorList>']
        compress = n % 2 == 0
        rows += ['          <binaryDataArrayList count="2">',
                 array(spectrum[:, 0].tolist(), 64, compress, "MS:1000514", "m/z array", MZ_UNIT, 12),
                 array(spectrum[:, 1].tolist(), 32, not compress, "MS:1000515", "intensity array", I_UNIT, 12),
                 '          </binaryDataArrayList>', '        </spectrum>']
        emit("\n".join(rows) + "\n")
    emit("      </spectrumList>\n")
    emit('      <chromatogramList count="1" defaultDataProcessingRef="pwiz_Reader_conversion">\n        ')
    chroma_offset = size
    emit('<chromatogram index="0" id="TIC" defaultArrayLength="%d">\n          <cvParam cvRef="MS" accession="MS:1000235" '
         'name="total ion current chromatogram" value=""/>\n          <binaryDataArrayList count="2">\n' % len(spectra))
    emit(array(rts, 64, True, "MS:1000595", "time array", T_UNIT, 12) + "\n")
    emit(array([float(s[:, 1].sum()) for s in spectra], 32, False, "MS:1000515", "intensity array", I_UNIT, 12) + "\n")
    emit("          </binaryDataArrayList>\n        </chromatogram>\n      </chromatogramList>\n    </run>\n  </mzML>\n")
    list_offset = size + 2
    emit('  <indexList count="2">\n    <index name="spectrum">\n')
    for n, at in enumerate(offsets):
        emit('      <offset idRef="controllerType=0 controllerNumber=1 scan=%d">%d</offset>\n' % (19 + n, at))
    emit('    </index>\n    <index name="chromatogram">\n      <offset idRef="TIC">%d</offset>\n    </index>\n  </indexList>\n' % chroma_offset)
    emit("  <indexListOffset>%d</indexListOffset>\n  <fileChecksum>0000000000000000000000000000000000000000</fileChecksum>\n</indexedmzML>\n" % list_offset)
    raw = "".join(out).encode("utf-8")
    # the index must point at the tags it names
    assert all(raw[at:at + 9] == b"<spectrum" for at in offsets) and raw[chroma_offset:chroma_offset + 13] == b"<chromatogram"
    assert raw[list_offset:list_offset + 10] == b"<indexList"
    expected = dict(peaks=np.concatenate(spectra), offsets=np.concatenate([[0], np.cumsum([len(s) for s in spectra])]).astype(np.int64),
                    levels=np.array(levels), rts=np.array(rts), ids=np.arange(19, 19 + len(spectra)))
    return raw, expected


def main():
    raw, expected = document(9)
    (HERE / "pwiz_style_small.mzML").write_bytes(raw)
    np.savez_compressed(HERE / "pwiz_style_small_expected.npz", **expected)
    print("wrote", len(raw), "bytes,", len(expected["ids"]), "spectra")


if __name__ == "__main__":
    main()
