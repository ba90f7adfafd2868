"""CLI mirroring the reference's `asr <model_dir> <audio_file> [language]` (/root/reference/src/main.rs:7-81)."""
import sys


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        print("Usage: python -m qwen3_asr_rs_b200 <model_dir> <audio_file> [language]", file=sys.stderr)   # main.rs:18-27
        return 1
    model_dir, audio, language = argv[0], argv[1], (argv[2] if len(argv) > 2 else None)
    from . import AsrInference
    eng = AsrInference.load(model_dir, device=0)
    try:
        r = eng.transcribe(audio, language)
    finally:
        eng.close()
    print(f"Language: {r.language}")                       # main.rs:77-78
    print(f"Text: {r.text}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
