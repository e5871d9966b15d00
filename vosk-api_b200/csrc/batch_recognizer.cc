// batch_recognizer.cc — see batch_recognizer.h.  Replaces [REF src/batch_recognizer.cc].
#include "batch_recognizer.h"

#include "vb_lattice.h"

#include <algorithm>
#include <cmath>

using namespace vb;

BatchRecognizer::BatchRecognizer(BatchModel *model, float sample_frequency)
    : model_(model), sample_frequency_(sample_frequency), sink_(std::make_shared<Sink>()),
      // [REF src/batch_recognizer.cc:27-29]
      resampler_(sample_frequency, 16000.0f, std::min(sample_frequency / 2, 16000.0f / 2), 6) {
    id_ = model->GetID(this);
    stream_ = model->engine_for(id_).open_stream();
    std::shared_ptr<Sink> sink = sink_;
    const Model *m = &model->model();
    const float lattice_beam = model->engine_for(id_).config().lattice_beam;
    const bool host_chain = model->engine_for(id_).config().lattice == 1;  // lattice=2: device lattice only, text from the best path
    stream_->on_result = [sink, m, lattice_beam, host_chain](const BestPath &bp) {
        // lattice=1: PushLattice's chain on the pruned raw lattice [REF src/batch_recognizer.cc:43-56]; otherwise (and if the
        // lattice came back empty or over capacity) the best path, which is the MBR result of a linear lattice
        std::vector<WordSpan> words;
        bool done = false;
        if (host_chain && bp.lattice && bp.lattice->error == 0 && bp.lattice->n_states > 0) {
            words = lattice_to_words(*bp.lattice, *m, lattice_beam);
            done = !words.empty() || bp.arcs.empty();
        }
        if (!done) words = align_words(*m, bp.arcs);
        std::lock_guard<std::mutex> lk(sink->mu);
        sink->early[bp.seq] = sink->nlsml ? result_nlsml(*m, words) : result_json(*m, words, bp.offset);
        for (auto it = sink->early.find(sink->next_seq); it != sink->early.end(); it = sink->early.find(sink->next_seq)) {
            sink->results.push(std::move(it->second));  // results of a stream come out in segment order
            sink->early.erase(it);
            sink->next_seq++;
        }
    };
}

BatchRecognizer::~BatchRecognizer() {}

void BatchRecognizer::EnableCapture() { stream_->capture.reset(new Capture); }

void BatchRecognizer::SetNLSML(bool nlsml) {
    std::lock_guard<std::mutex> lk(sink_->mu);
    sink_->nlsml = nlsml;
}

void BatchRecognizer::AcceptWaveform(const char *data, int len) {
    if (finished_ || len < 2) return;
    const int n = len / 2;  // odd trailing byte ignored, as in the reference (len / 2)
    const int16_t *pcm = reinterpret_cast<const int16_t *>(data);
    if (resampler_.identity()) {
        buffer_.insert(buffer_.end(), pcm, pcm + n);
    } else {
        std::vector<float> in(n), out;
        for (int i = 0; i < n; i++) in[i] = pcm[i];
        resampler_.resample_flush(in, &out);
        for (float v : out) buffer_.push_back((int16_t)std::lrintf(std::max(-32768.f, std::min(32767.f, v))));
    }
    const int spc = model_->samples_per_chunk();
    size_t i = 0;
    Engine &eng = model_->engine_for(id_);
    while (i + spc <= buffer_.size()) {
        eng.push(stream_, buffer_.data() + i, spc, false);
        i += spc;
    }
    if (i) buffer_.erase(buffer_.begin(), buffer_.begin() + i);
}

void BatchRecognizer::FinishStream() {
    if (finished_) return;
    finished_ = true;
    // whatever is buffered (possibly nothing) goes out flagged last [REF src/batch_recognizer.cc:37-41]
    model_->engine_for(id_).push(stream_, buffer_.data(), (int)buffer_.size(), true);
    buffer_.clear();
}

const char *BatchRecognizer::FrontResult() {
    std::lock_guard<std::mutex> lk(sink_->mu);
    if (sink_->results.empty()) return "";
    front_ = sink_->results.front();
    return front_.c_str();
}

void BatchRecognizer::Pop() {
    std::lock_guard<std::mutex> lk(sink_->mu);
    if (!sink_->results.empty()) sink_->results.pop();
}

const char *BatchRecognizer::PartialResult() {
    std::vector<WordSpan> words;
    {
        std::lock_guard<std::mutex> lk(stream_->partial_mu);
        for (int w : stream_->partial_words) words.push_back(WordSpan{w, 0.f, 0.f, 1.f});
    }
    partial_ = partial_json(model_->model(), words);
    return partial_.c_str();
}

int BatchRecognizer::PartialFrames() {
    std::lock_guard<std::mutex> lk(stream_->partial_mu);
    return stream_->partial_frames;
}

int BatchRecognizer::GetNumPendingChunks() { return stream_->pending_chunks.load(); }
