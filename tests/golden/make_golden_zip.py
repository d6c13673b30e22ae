"""Expectations for the reference's ZIP fixtures (test/_data/zip/, used by test/zip_test.dart:1-211 and :731-775):
member names, sizes, SHA-256 of the contents, CRCs and modes as CPython's zipfile reads them (an independent reader).
Run in the build container (needs /root/reference only to copy the fixtures); writes tests/golden/zip/manifest.json."""
import hashlib, json, os, zipfile
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zip")
man = {}
for f in sorted(os.listdir(HERE)):
    if not (f.endswith(".zip") or f.endswith(".notzip")):
        continue
    p = os.path.join(HERE, f)
    try:
        z = zipfile.ZipFile(p)
    except zipfile.BadZipFile:
        man[f] = {"bad": True}
        continue
    ents = []
    for i in z.infolist():
        data = z.read(i) if not i.flag_bits & 1 else None
        ents.append({"name": i.filename, "size": i.file_size, "crc32": i.CRC, "method": i.compress_type, "mode": i.external_attr >> 16,
                     "sha256": hashlib.sha256(data).hexdigest() if data is not None else None})
    man[f] = {"entries": ents}
json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
print({k: (len(v.get("entries", [])) if "entries" in v else "bad") for k, v in man.items()})
