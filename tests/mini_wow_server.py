"""TEST INFRASTRUCTURE — a minimal restatement of the reference server's WOW request path, so the drop-in modules can be
driven through FastAPI on the GPU box (where /root/reference does not exist).  It follows server/app/main.py:

* ``POST /api/wow`` (main.py:458-545): resolves ``input_file``, creates ``job_id = wow_<timestamp>``, registers the job in
  ``sr_jobs`` with status "queued" and schedules ``run_wow_job`` as a Starlette background task; answers ``SRResponse``.
* ``run_wow_job`` (main.py:290-368): status "processing" -> ``process_wow_sr(input_tif, output_dir, enhance_crops, model)`` ->
  (tiling, only when a GeoTIFF was written) -> status "completed" with ``result``; any exception -> status "failed".
* ``POST /api/enhance`` (main.py:545-640): multipart upload, model validation (400), job registered, ``run_wow_job`` with
  ``enhance_crops=True``.
* ``GET /api/sr/{job_id}`` (main.py:437-443).

``process_wow_sr`` is whatever module is passed in: the drop-in (``wowsr_b200.app.wow_sr``) in the GPU test."""
from __future__ import annotations

import itertools
from datetime import datetime
from pathlib import Path
from typing import Optional

from fastapi import BackgroundTasks, FastAPI, File, Form, HTTPException, UploadFile
from pydantic import BaseModel


class WowRequest(BaseModel):          # main.py:200-208
    input_file: Optional[str] = None
    enhance_crops: bool = True
    auto_fetch: bool = True
    max_age_days: int = 30
    max_cloud_cover: float = 30.0
    force_fetch: bool = False


class SRResponse(BaseModel):          # main.py:~190
    job_id: str
    status: str
    message: str


def build_app(wow_sr_module, data_dir: Path) -> FastAPI:
    app = FastAPI()
    sr_jobs: dict = {}
    app.state.sr_jobs = sr_jobs
    counter = itertools.count()

    def run_wow_job(job_id, input_file, output_dir, enhance_crops, model="realesrgan_x4"):
        try:
            sr_jobs[job_id]["status"] = "processing"
            result = wow_sr_module.process_wow_sr(input_tif=input_file, output_dir=output_dir, enhance_crops=enhance_crops, model=model)
            sr_jobs[job_id]["status"] = "completed"
            sr_jobs[job_id]["message"] = "WOW Super-resolution complete!"
            sr_jobs[job_id]["result"] = result
        except Exception as e:  # noqa: BLE001  (main.py:365-368)
            sr_jobs[job_id]["status"] = "failed"
            sr_jobs[job_id]["message"] = str(e)

    @app.post("/api/wow", response_model=SRResponse)
    async def start_wow_sr(request: WowRequest, background_tasks: BackgroundTasks):
        if not request.input_file:
            raise HTTPException(status_code=404, detail="No GeoTIFF files found. Enable auto_fetch=true or run fetch first.")
        input_file = Path(request.input_file)
        if not input_file.exists():
            raise HTTPException(status_code=404, detail=f"Input file not found: {input_file}")
        job_id = f"wow_{datetime.now().strftime('%Y%m%d_%H%M%S')}_{next(counter)}"
        output_dir = data_dir / "wow" / job_id
        output_dir.mkdir(parents=True, exist_ok=True)
        sr_jobs[job_id] = {"status": "queued", "message": "WOW job queued (Real-ESRGAN x4 + Enhanced)", "input_file": str(input_file),
                           "pipeline": "RealESRGAN_x4 + Enhanced", "scale": 4, "enhance_crops": request.enhance_crops,
                           "output_dir": str(output_dir), "created_at": datetime.now().isoformat()}
        background_tasks.add_task(run_wow_job, job_id, input_file, output_dir, request.enhance_crops)
        return SRResponse(job_id=job_id, status="queued", message=f"WOW SR started: {input_file.name}")

    @app.post("/api/enhance")
    async def enhance_image_upload(image: UploadFile = File(...), model: str = Form("realesrgan_x4"), background_tasks: BackgroundTasks = None):
        valid_models = ["realesrgan_x4", "realesrgan_anime"]
        if model not in valid_models:
            raise HTTPException(status_code=400, detail=f"Invalid model. Choose from: {valid_models}")
        content = await image.read()
        job_id = f"wow_{datetime.now().strftime('%Y%m%d_%H%M%S')}_{next(counter)}"
        output_dir = data_dir / "wow" / job_id
        upload_dir = data_dir / "uploads" / job_id
        output_dir.mkdir(parents=True, exist_ok=True)
        upload_dir.mkdir(parents=True, exist_ok=True)
        uploaded_path = upload_dir / image.filename
        uploaded_path.write_bytes(content)
        sr_jobs[job_id] = {"status": "processing", "message": "Enhancement starting", "input_file": str(uploaded_path),
                           "output_dir": str(output_dir), "model": model, "created_at": datetime.now().isoformat()}
        background_tasks.add_task(run_wow_job, job_id, uploaded_path, output_dir, True, model)
        return {"job_id": job_id, "status": sr_jobs[job_id]["status"], "message": sr_jobs[job_id]["message"], "model": model}

    @app.get("/api/sr/{job_id}")
    async def get_sr_status(job_id: str):
        if job_id not in sr_jobs:
            raise HTTPException(status_code=404, detail="Job not found")
        return sr_jobs[job_id]

    return app
