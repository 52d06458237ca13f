"""``/generate?text=...&speaker_id=N`` → ``audio/wav`` with an md5 cache — the serving endpoint of the reference's ``app.py``
(:55-99,121-134) on the standard library's HTTP server (Flask is not available here; the HTML / JS demo page under ``web/`` is
out of scope).

    python app.py --load_path logs/<run> --num_speakers 2 --port 5000
    curl 'http://localhost:5000/generate?text=안녕하세요&speaker_id=0' -o out.wav

As in the reference a request is answered from ``<audio_root>/<model name>/<md5(text)>.<speaker_id>.0.wav`` when that file
exists, otherwise the utterance is synthesised (attention-trimmed), written there and sent; a failed synthesis answers
``{"success": false}`` with status 400.  ``tokens=5 9 23 1`` may replace ``text=`` (symbol ids ending in EOS).  One engine,
one request at a time (a lock serialises synthesis; cached files are served concurrently).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from urllib.parse import parse_qs, urlparse


def audio_cache_path(audio_root: str, model_name: str, text: str, speaker_id: int) -> str:
    """app.py:58-65: md5 of the UTF-8 text, one directory per model, the ``.0`` of synthesize()'s per-utterance postfix."""
    hashed = hashlib.md5(text.encode("utf-8")).hexdigest()
    return os.path.join(audio_root, model_name, "{}.{}.0.wav".format(hashed, speaker_id))


def make_server(synthesizer, load_path: str, port: int = 5000, audio_root: str = os.path.join("web", "audio"),
                host: str = "0.0.0.0") -> ThreadingHTTPServer:
    model_name = os.path.basename(os.path.normpath(load_path))
    lock = threading.Lock()

    class Handler(BaseHTTPRequestHandler):
        def log_message(self, fmt, *args):                      # quiet by default, like Flask with debug off
            pass

        def _json(self, status, payload):
            body = json.dumps(payload).encode()
            self.send_response(status)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.send_header("Access-Control-Allow-Origin", "*")    # flask_cors.CORS(app), app.py:23
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            url = urlparse(self.path)
            q = parse_qs(url.query)
            if url.path != "/generate":
                return self._json(404, {"success": False, "error": "only /generate is served (the demo page is not part of this build)"})
            text = (q.get("text") or [None])[0]
            tokens = (q.get("tokens") or [None])[0]
            try:
                speaker_id = int((q.get("speaker_id") or ["0"])[0])
            except ValueError:
                return self._json(400, {"success": False})
            if not text and not tokens:
                return self._json(200, {})                      # app.py:96-99
            key = text if text else "tokens:" + " ".join(tokens.split())
            path = audio_cache_path(audio_root, model_name, key, speaker_id)
            if not os.path.exists(path):
                try:
                    kw = dict(texts=[text]) if text else dict(tokens=[[int(t) for t in tokens.split()]])
                    with lock:
                        if not os.path.exists(path):
                            tmp = path + ".part"
                            synthesizer.synthesize(paths=[tmp], speaker_ids=[speaker_id], attention_trim=True, **kw)
                            os.replace(tmp, path)
                except Exception:
                    return self._json(400, {"success": False})  # app.py:73-74
            with open(path, "rb") as f:
                body = f.read()
            self.send_response(200)
            self.send_header("Content-Type", "audio/wav")
            self.send_header("Content-Disposition", 'attachment; filename="{}.wav"'.format(os.path.basename(path).split(".")[0]))
            self.send_header("Content-Length", str(len(body)))
            self.send_header("Access-Control-Allow-Origin", "*")
            self.end_headers()
            self.wfile.write(body)

    return ThreadingHTTPServer((host, port), Handler)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--load_path", required=True)
    parser.add_argument("--checkpoint_step", default=None, type=int)
    parser.add_argument("--num_speakers", default=1, type=int)
    parser.add_argument("--port", default=5000, type=int)
    parser.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16"])
    config = parser.parse_args(argv)
    if not os.path.exists(config.load_path):
        print(" [!] load_path not found: {}".format(config.load_path))
        return 1
    from .synthesizer import Synthesizer
    from .text import text_to_sequence
    synthesizer = Synthesizer(precision=config.precision, text_to_sequence=text_to_sequence)
    synthesizer.load(config.load_path, config.num_speakers, config.checkpoint_step)
    server = make_server(synthesizer, config.load_path, config.port)
    try:
        server.serve_forever()
    finally:
        server.server_close()
        synthesizer.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
