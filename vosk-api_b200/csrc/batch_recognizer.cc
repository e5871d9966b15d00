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
    // non-16 kHz input is resampled on the GPU (option device-resample, default on) for rates up to 48 kHz
    device_resample_ = !resampler_.identity() && model->engine_for(id_).config().device_resample && sample_frequency <= 48000.f;
    std::shared_ptr<Sink> sink = sink_;
    const Model *m = &model->model();
    const float lattice_beam = model->engine_for(id_).config().lattice_beam;
    const bool host_chain = model->engine_for(id_).config().lattice == 1;  // lattice=2: device lattice only, text from the best path
    Engine *eng = &model->engine_for(id_);
    stream_->on_result = [sink, m, lattice_beam, host_chain, eng](const BestPath &bp) {
        // lattice=1: PushLattice's chain on the pruned raw lattice [REF src/batch_recognizer.cc:43-56]; otherwise (and if the
        // lattice came back empty or over capacity) the best path, which is the MBR result of a linear lattice
        std::vector<WordSpan> words;
        bool done = false;
        if (host_chain) {
            const RawLattice *lat = bp.raw_lattice();
            bool ran = false;
            if (lat && lat->error == 0 && lat->n_states > 0) words = lattice_to_words(*lat, *m, lattice_beam, 0.9, nullptr, &ran);
            done = ran || bp.arcs.empty();
            if (!done) eng->count_fallback();  // no usable lattice (capacity error, no complete path): logged by the engine, counted here
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

BatchRecognizer::~BatchRecognizer() {
    // a stream that was never finished still owns an engine channel: close it with an empty last chunk (whatever was buffered
    // is dropped; the result lands in the shared sink, which nobody reads any more)
    if (!finished_) {
        finished_ = true;
        try {
            Stream::Chunk ck;
            ck.last = true;
            model_->engine_for(id_).push_chunk(stream_, std::move(ck));
        } catch (...) {
        }
    }
}

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
    } else if (device_resample_) {
        // keep the call as it came; hand over chunks of samples_per_chunk OUTPUT samples with the raw samples their taps reach
        Call c;
        c.in.assign(pcm, pcm + n);
        c.n_out = resampler_.num_output(n);
        avail_ += c.n_out;
        if (c.n_out > 0) calls_.push_back(std::move(c));
        const int spc = model_->samples_per_chunk();
        while (avail_ >= spc) push_resampled_chunk(spc, false);
        return;
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
    if (device_resample_) {
        push_resampled_chunk((int)avail_, true);
        return;
    }
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

// One chunk of n_out 16 kHz samples, described by the calls it spans: per call the output range taken and the raw
// samples those outputs' filter taps can reach (the reference resamples every call on its own with flush=true
// [REF src/batch_recognizer.cc:157-158], so taps never cross a call boundary).  A chunk made of too many small calls
// is resampled here instead (same arithmetic, vb_result.cc) and handed over as plain 16 kHz samples.
void BatchRecognizer::push_resampled_chunk(int n_out, bool last) {
    Engine &eng = model_->engine_for(id_);
    Stream::Chunk ck;
    ck.last = last;
    ck.rate = resampler_.in_rate();
    ck.n_out = n_out;
    int pos = 0;
    size_t k = 0;
    bool fits = true;
    for (; pos < n_out; k++) {
        Call &c = calls_[k];
        const long long take = std::min<long long>(c.n_out - c.taken, n_out - pos);
        long long lo, hi;
        resampler_.input_range(c.taken, c.taken + take - 1, &lo, &hi);
        lo = std::max<long long>(lo, 0);
        hi = std::min<long long>(hi, (long long)c.in.size() - 1);
        Stream::Chunk::Seg g{(int)ck.raw.size(), (int)lo, (int)c.in.size(), (int)c.taken, pos, (int)take};
        if (hi >= lo) ck.raw.insert(ck.raw.end(), c.in.begin() + lo, c.in.begin() + hi + 1);
        ck.segs.push_back(g);
        pos += (int)take;
        if ((int)ck.segs.size() > Engine::kMaxResampleSegs || (int)ck.raw.size() > eng.max_resample_raw()) fits = false;
    }
    if (n_out == 0) ck.rate = 0;  // an empty last chunk needs no device pass
    if (!fits) {  // host path for this chunk
        ck.samples.resize(n_out);
        int p2 = 0;
        for (size_t j = 0; p2 < n_out; j++) {
            Call &c = calls_[j];
            const long long take = std::min<long long>(c.n_out - c.taken, n_out - p2);
            std::vector<float> in(c.in.begin(), c.in.end()), out;
            resampler_.resample_flush(in, &out);
            for (long long t = 0; t < take; t++)
                ck.samples[p2 + t] = (int16_t)std::lrintf(std::max(-32768.f, std::min(32767.f, out[(size_t)(c.taken + t)])));
            p2 += (int)take;
        }
        ck.rate = 0;
        ck.n_out = 0;
        ck.raw.clear();
        ck.segs.clear();
    }
    // consume
    int left = n_out;
    while (left > 0) {
        Call &c = calls_.front();
        const long long take = std::min<long long>(c.n_out - c.taken, left);
        c.taken += take;
        left -= (int)take;
        if (c.taken == c.n_out) calls_.pop_front();
    }
    avail_ -= n_out;
    eng.push_chunk(stream_, std::move(ck));
}
