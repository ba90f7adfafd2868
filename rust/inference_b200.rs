// The `#[cfg(feature = "b200")]` arm of `src/inference.rs`: same public surface (`AsrInference::{load, transcribe}`,
// `TranscribeResult`), steps 1 and 9 untouched, steps 2-8 (`src/inference.rs:94-200`) replaced by one library call.
#[cfg(feature = "b200")]
mod b200_arm {
    use crate::audio;
    use crate::backend::b200::B200Engine;
    use crate::mel::MEL_SAMPLE_RATE;
    use crate::tokenizer::AsrTokenizer;
    use crate::{capitalize_first, parse_asr_output, TranscribeResult};
    use anyhow::Result;

    pub struct AsrInference {
        engine: B200Engine,
        tokenizer: AsrTokenizer,
    }

    impl AsrInference {
        /// `src/inference.rs:30-86`; the device pick of `src/main.rs:51-58` becomes the `device` ordinal.
        pub fn load(model_dir: &str, device: i32) -> Result<Self> {
            let tokenizer = AsrTokenizer::from_dir(model_dir)?;                      // unchanged (:76-84)
            let engine = B200Engine::load(model_dir, device, 4096)?;                 // cap 4096 new tokens (:153); session sized per clip
            Ok(Self { engine, tokenizer })
        }

        /// `src/inference.rs:89-213`
        pub fn transcribe(&self, audio_path: &str, language: Option<&str>) -> Result<TranscribeResult> {
            let samples = audio::load_audio(audio_path, MEL_SAMPLE_RATE)?;           // step 1, unchanged (:92)
            let lang_ids: Option<Vec<i64>> = match language {                         // :246-250: no <asr_text> suffix
                Some(l) => Some(self.tokenizer.encode(&format!("language {}", capitalize_first(l)))?),
                None => None,
            };
            let generated = self.engine.transcribe_ids(&samples, lang_ids.as_deref())?;   // steps 2-8
            let raw_text = self.tokenizer.decode(&generated)?;                        // step 9, unchanged (:204-206)
            let (language_detected, transcription) = parse_asr_output(&raw_text, language.is_some());
            Ok(TranscribeResult { text: transcription, language: language_detected, raw_output: raw_text })
        }
    }
}
